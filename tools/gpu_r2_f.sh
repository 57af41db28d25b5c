#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/ab_time.py head=ab/lib_head.so,nodrain=ab/lib_nodrain.so,drain=default "16384,16384,512,14;16384,16384,1024,14;16384,16384,2048,14;16384,16384,16384,14;4096,4096,4096,14" 2 > gpurun_out/r02_ab_f.jsonl 2> gpurun_out/r02_ab_f.err
cat gpurun_out/r02_ab_f.jsonl; tail -3 gpurun_out/r02_ab_f.err
