import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import rapidobjectdetectionusingcascadedcnns_b200 as pkg
from rapidobjectdetectionusingcascadedcnns_b200 import synth
B=32
frames=[synth.synth_frame(480,640,j) for j in range(B)]
app=pkg.InferenceCascadeApp("p", stage_models=[pkg.StageModel(w) for w in synth.xavier_cascade()])
infos=[pkg.ArrayImageInfo(f,"f%d.png"%j) for j,f in enumerate(frames)]
for _ in range(3): app.run_inference_on_images(infos)
eng=app._engine
def t(fn,n=10):
    torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): r=fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
print("full api        %.2f ms" % t(lambda: app.run_inference_on_images(infos)))
d=torch.from_numpy(np.stack(frames)).cuda()
print("engine resident %.2f ms" % t(lambda: eng.run_frames(d)))
print("stage_host      %.2f ms" % t(lambda: eng.stage_host_frames(frames)))
pin=eng.stage_host_frames(frames)
print("h2d             %.2f ms" % t(lambda: pin.to('cuda', non_blocking=True)))
print("engine from pin %.2f ms" % t(lambda: eng.run_frames(pin)))
print("upload (gather + h2d pipelined) %.2f ms" % t(lambda: eng.upload_host_frames(frames)))
print("engine from upload %.2f ms" % t(lambda: eng.run_frames(eng.upload_host_frames(frames))))
res=eng.run_frames(d)
print("boxes           %.2f ms  (%d boxes)" % (t(lambda: app._boxes_from_result(res, infos, B)), res.boxes.shape[0]))
print("load_frame x32  %.2f ms" % t(lambda: [app._load_frame(i) for i in infos]))
