#!/bin/bash
cd "$(dirname "$0")/.."; mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for t in memcheck racecheck synccheck; do echo "== compute-sanitizer $t"; timeout 500 compute-sanitizer --tool $t python tools/sanitize_run.py > gpurun_out/sanitizer_$t.log 2>&1; tail -3 gpurun_out/sanitizer_$t.log; done
for x in MMult_cuda_9 MMult_cuda_11; do f=gpurun_out/output_ref_harness_$x.m; echo "version = '$x';" > $f; timeout 300 oracle/_ref/ref_cuda_test_MMult__$x.x >> $f 2>&1; tail -3 $f; done
python - <<PY
import sys, torch, statistics
sys.path.insert(0,"tests"); import _libs
g=_libs.load_pkg()
N=8192
A=(torch.rand(N,N,device="cuda")-0.5).bfloat16(); B=(torch.rand(N,N,device="cuda")-0.5).bfloat16(); C=torch.empty(N,N,device="cuda",dtype=torch.bfloat16)
res={}
for trial in range(4):
  for rows in (512,1024,2048,4096,8192):
    g.lib.b200_gemm_debug_set_group_rows(rows)
    for _ in range(2): g.gemm_bf16(A,B,out=C)
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True); s.record()
    for _ in range(10): g.gemm_bf16(A,B,out=C)
    e.record(); torch.cuda.synchronize(); res.setdefault(rows,[]).append(s.elapsed_time(e)/10)
for rows,v in res.items(): print("raster rows %5d: best %.0f med %.0f TF"%(rows, 2*N**3/min(v)/1e9, 2*N**3/statistics.median(v)/1e9))
g.lib.b200_gemm_debug_set_group_rows(0)
for N in (256,512,1024,2048,4096):
    A=torch.rand(N,N,device="cuda")-0.5; B=torch.rand(N,N,device="cuda")-0.5; C=torch.empty(N,N,device="cuda")
    for md in (4,2,5,3,0,1):
        for _ in range(3): g.gemm_f32(A,B,out=C,mode=md)
        torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True); s.record()
        for _ in range(20): g.gemm_f32(A,B,out=C,mode=md)
        e.record(); torch.cuda.synchronize(); ms=s.elapsed_time(e)/20
        print("N %d mode %d %s us %.1f TF %.1f"%(N, md, g.last_kernel(), ms*1e3, 2*N**3/ms/1e9))
PY
