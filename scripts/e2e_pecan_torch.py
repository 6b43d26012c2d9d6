#!/usr/bin/env python
"""Development aid: scripts/e2e_pecan.py with torch imported first (does the import change the host-side phases?)."""
import os, sys, runpy
import torch  # noqa: F401
torch.cuda.set_device(0)
print("torch threads", torch.get_num_threads(), "OMP_NUM_THREADS", os.environ.get("OMP_NUM_THREADS"), flush=True)
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "e2e_pecan.py"), run_name="__main__")
