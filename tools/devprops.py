#!/usr/bin/env python
"""Prints the device properties the design depends on (L2 size, persisting carve-out, policy window, smem)."""
import torch
p = torch.cuda.get_device_properties(0)
print(p)
from cuda import cudart
err, prop = cudart.cudaGetDeviceProperties(0)
for k in ("l2CacheSize", "persistingL2CacheMaxSize", "accessPolicyMaxWindowSize", "sharedMemPerMultiprocessor",
          "sharedMemPerBlockOptin", "regsPerMultiprocessor", "multiProcessorCount", "clockRate", "memoryClockRate", "memoryBusWidth"):
    print(k, getattr(prop, k, None))
