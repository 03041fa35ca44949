#!/bin/bash
# bring-up of the TMA conv path: one launch per debug variant (VMB_CONV_DBG bits: 1 prefetch.tensormap, 2 no swizzle, 4 skip weights box, 8 skip input box)
for d in 0 1 2 4 8 12; do
  CUDA_LAUNCH_BLOCKING=1 VMB_CONV_DBG=$d timeout 60 python tools/conv_one.py down1_2 up2_1 > /dev/null 2> gpurun_out/dbg_$d.err && echo "DBG=$d ok" || { echo "DBG=$d FAIL"; tail -2 gpurun_out/dbg_$d.err; }
done
VMB_CONV_DBG=0 timeout 120 compute-sanitizer --tool memcheck python tools/conv_one.py down1_2 2>&1 | grep -v "^$" | head -30
