"""Diagnostic: does importing torch / pinning memory in the same process slow the copy engine?"""
import os, sys, time
mode = sys.argv[1]
if mode in ("torch", "torch_pin", "torch_cuda"):
    import torch
    if mode == "torch_cuda":
        torch.cuda.set_device(0); torch.zeros(1, device="cuda")
    if mode == "torch_pin":
        t0 = time.time(); keep = torch.empty((138569, 5408), dtype=torch.float64, pin_memory=True); print("pin %.2fs" % (time.time() - t0))
sys.argv = [sys.argv[0], "3"]
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "e2e_quick.py")).read())
