"""Which side of a 2-D D2H copy costs the 9 % against a contiguous copy: the pitched device source, the pitched host destination, or the
row length? cuda-python runtime calls, pinned host memory, CUDA-event timing."""
import json

import torch
from cuda import cudart

N = 20000
dev = torch.device("cuda", 0)
Kd = torch.zeros((N, N), dtype=torch.float64, device=dev)
Kp = torch.zeros((N, N), dtype=torch.float64).pin_memory()
st = torch.cuda.Stream()
D2H = cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost


def timed(fn, reps=3):
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(st):
            e0.record()
            fn()
            e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def copy2d(dpitch, spitch, width, rows):
    (err,) = cudart.cudaMemcpy2DAsync(Kp.data_ptr(), dpitch, Kd.data_ptr(), spitch, width, rows, D2H, st.cuda_stream)
    assert err == cudart.cudaError_t.cudaSuccess, err


out = {}
for name, dp, sp, w, r in (
    ("rect_19000_cols_both_pitched", N * 8, N * 8, 19000 * 8, 10000),
    ("rect_19000_cols_host_dense", 19000 * 8, N * 8, 19000 * 8, 10000),
    ("rect_19000_cols_dev_dense", N * 8, 19000 * 8, 19000 * 8, 10000),
    ("rect_19000_cols_both_dense", 19000 * 8, 19000 * 8, 19000 * 8, 10000),
    ("rect_2000_cols_both_pitched", N * 8, N * 8, 2000 * 8, 20000),
    ("rect_2000_cols_host_dense", 2000 * 8, N * 8, 2000 * 8, 20000),
    ("rect_512_cols_both_pitched", N * 8, N * 8, 512 * 8, 20000),
    ("rect_10000_cols_both_pitched", N * 8, N * 8, 10000 * 8, 20000),
    ("rect_10000_cols_pitch_plus_64B", 10008 * 8, 10008 * 8, 10000 * 8, 20000),
):
    ms = timed(lambda: copy2d(dp, sp, w, r))
    out[name] = {"ms": round(ms, 3), "GBps": round(w * r / ms / 1e6, 2)}
    print(name, out[name], flush=True)
print(json.dumps(out))
