"""In-memory logger with the reference's interface (`utils/log.py:22-62`)."""
import time

log_cache = []
console_output_enabled = False


def log(msg: str, log_to_console=True, log_to_file=True):
    line = time.strftime('%X') + ": {}".format(msg)
    if log_to_console and console_output_enabled:
        print(line)
    if log_to_file:
        log_cache.append(time.strftime('%x') + ' ' + line)
        if len(log_cache) > 100000:
            del log_cache[:50000]


def log_clear():
    del log_cache[:]


def enable_console_output():
    global console_output_enabled
    console_output_enabled = True


def disable_console_output():
    global console_output_enabled
    console_output_enabled = False
