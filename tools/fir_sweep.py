"""Tensor-core vs FP32 FirFilter (real taps, c32) over tap/decimation shapes: which path the planner should take.
   python tools/fir_sweep.py  (GPU box)"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rustradio_b200 as R
from oracle import oracle as O

def timeit(f, din, need, dout, out_n, reps=10):
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        f.run(din, need, dout, out_n, st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        f.run(din, need, dout, out_n, st)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

F32 = "--f32" in sys.argv                      # FirFilter<Float>: f32 samples, twice as many for the same bytes
n = 1 << (26 if F32 else 25)
x = torch.empty(n if F32 else 2 * n, dtype=torch.float32, device="cuda")
R.synth_f32(x, 7, 0, n if F32 else 2 * n, 0, torch.cuda.current_stream().cuda_stream)
shapes = [(16, 1), (24, 1), (32, 2), (48, 2), (64, 4), (96, 4), (32, 1), (64, 1), (121, 1), (128, 1), (247, 1), (512, 1), (1025, 1), (64, 2), (127, 2), (255, 2), (129, 4), (255, 4), (511, 4),
          (255, 8), (511, 8), (255, 10), (1023, 16)]
rows = []
for T, D in shapes:
    taps = O.low_pass_n(1.0, 0.4 / D, T)
    if not F32:
        taps = taps.astype(np.complex64)
        if "--ctaps" in sys.argv:                      # complex taps (what translate() produces)
            taps = (taps * np.exp(2j * np.pi * 0.05 * np.arange(T))).astype(np.complex64)
    res = {}
    for name, env in (("fp32", "0"), ("tensor", "2")):
        os.environ["RRC_FIR_TENSOR"] = env
        f = R.Fir(taps, deci=D)
        out_n = f.out_count(n)
        need = (out_n - 1) * D + T
        y = torch.empty(out_n if F32 else 2 * out_n, dtype=torch.float32, device="cuda")
        if name == "tensor" and not f.uses_tensor_cores:
            res[name] = None
            continue
        res[name] = timeit(f, x, need, y, out_n)
    os.environ["RRC_FIR_TENSOR"] = "1"
    auto = R.Fir(taps, deci=D).uses_tensor_cores
    hbm_ms = (4 if F32 else 8) * (n + n // D) / 6.45e12 * 1e3
    rows.append(dict(ntaps=T, deci=D, fp32_ms=res["fp32"], tensor_ms=res["tensor"], planner_takes_tensor=auto, hbm_floor_ms=hbm_ms))
    t = res["tensor"]
    print(f"T={T:5d} D={D:3d}  fp32 {res['fp32']:.3f} ms  tensor {t if t is None else round(t, 3)} ms  planner={'tensor' if auto else 'fp32'}  hbm floor {hbm_ms:.3f}", flush=True)
json.dump(rows, open("gpurun_out/fir_sweep_f32.json" if F32 else "gpurun_out/fir_sweep_ctaps.json" if "--ctaps" in sys.argv else "gpurun_out/fir_sweep.json", "w"), indent=1)
