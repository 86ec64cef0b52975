"""SAP front end per 4096 x 4096 x 3 uint8 image: device blur + Canny (ucf_gaussian_blur_u8 / ucf_canny_u8, CUDA events)
against OpenCV on the box's host cores.   python scripts/gpu_front_end_bench.py [n]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from ucf_vit_b200 import ops
from test_gpu_canny import _scene
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
img = _scene(n, np.random.default_rng(n))
x = torch.as_tensor(img).cuda()
def ev(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): out = fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, out
for k, lo in ((5, 60), (3, 80), (1, 50)):
    tb, b = ev(lambda: ops.gaussian_blur_u8(x, k))
    tc, (e, sweeps) = ev(lambda: ops.canny_u8(b, lo, lo + 50, return_sweeps=True))
    nbytes = img.size
    line = (f"{n}x{n}x3 k={k} low={lo}: device blur {tb*1e3:.0f} us ({2*nbytes/tb/1e6:.0f} GB/s of 1 read + 1 write), "
            f"canny {tc*1e3:.0f} us ({sweeps} hysteresis sweeps, each a host round trip), edges {100*(e>0).float().mean().item():.2f} %")
    try:
        import cv2
        t0 = time.perf_counter(); hb = cv2.GaussianBlur(img, (k, k), 0); t1 = time.perf_counter(); he = cv2.Canny(hb, lo, lo + 50); t2 = time.perf_counter()
        same = np.array_equal(he, e.cpu().numpy()) and np.array_equal(hb, b.cpu().numpy())
        line += f" | OpenCV {cv2.__version__} on {cv2.getNumThreads()} host threads: blur {1e3*(t1-t0):.1f} ms, canny {1e3*(t2-t1):.1f} ms; identical bytes: {same}"
    except ImportError:
        line += " | cv2 not importable"
    print(line)
t0 = time.perf_counter(); e_host = e.cpu(); t1 = time.perf_counter()
print(f"edge map D2H ({e.numel()/1e6:.1f} MB, pageable): {1e3*(t1-t0):.2f} ms")
