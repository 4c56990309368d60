"""Driver of the kind::f16 TS-mode micro-experiment (GPU box): python tools/exp_f16ts.py
   Build here (no GPU needed): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC tools/exp_f16ts.cu -o tools/exp_f16ts.so"""
import ctypes, os, sys
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "exp_f16ts.so"))
lib.exp_f16ts_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
for N in (64, 128):
    A = (torch.randint(-8, 9, (128, 64), generator=g).float() / 4).half().to(dev)
    B = (torch.randint(-8, 9, (N, 64), generator=g).float() / 8).half().to(dev)
    ref = A.float() @ B.float().t()
    for pack in (0, 1):
        D = torch.zeros((128, N), device=dev)
        rc = lib.exp_f16ts_run(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, pack, 0, 0, None)
        err = float((D - ref).abs().max())
        print(f"N={N} pack_mode={pack} ({'even k in the LOW half' if pack == 0 else 'even k in the HIGH half'}): rc={rc} max|D - A.B^T| = {err:.4g}  (max|ref| {float(ref.abs().max()):.3g})", flush=True)
cyc = torch.zeros(2, dtype=torch.int64, device=dev)
A = torch.zeros((128, 64), dtype=torch.half, device=dev); 
for N in (32, 64, 128):
    B = torch.zeros((N, 64), dtype=torch.half, device=dev)
    D = torch.zeros((128, N), device=dev)
    for mode, name in ((1, "kind::f16  (K=16)"), (2, "kind::tf32 (K=8) ")):
        for iters in (512, 4096):
            lib.exp_f16ts_run(A.data_ptr(), B.data_ptr(), D.data_ptr(), N, 0, mode, iters, cyc.data_ptr())
            c = cyc.cpu()
            print(f"N={N:3d} {name} iters={iters}: issue {int(c[0]) / iters:.1f} cycles/MMA, issue+drain {(int(c[0]) + int(c[1])) / iters:.1f} cycles/MMA", flush=True)
