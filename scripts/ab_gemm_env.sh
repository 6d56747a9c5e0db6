#!/bin/bash
# A/B of the GEMM cluster-multicast path via PI3B_GEMM_CLUSTER on the same box
cd scripts
for rep in 1 2; do
for v in 0 1; do
  echo "== PI3B_GEMM_CLUSTER=$v (rep $rep)"
  PI3B_GEMM_CLUSTER=$v timeout -s KILL 100 python -c "
import torch, sys
sys.path.insert(0,'..')
from pi3_slam_b200 import ops
from bench_kernels import time_cuda
dev='cuda'
for (M,N,K,act,res) in [(100400,3072,1024,0,0),(100400,1024,1024,0,1),(100400,4096,1024,1,0),(100400,1024,4096,0,1),(64300,3072,1024,0,0)]:
    a=torch.randn(M,K,device=dev).bfloat16(); w=(torch.randn(N,K,device=dev)*0.02).bfloat16(); b=torch.randn(N,device=dev); g=torch.rand(N,device=dev)
    if res:
        x=torch.randn(M,N,device=dev); f=lambda: ops.gemm(a,w,x,bias=b,gamma=g,resid=x)
    else:
        o=torch.empty(M,N,device=dev,dtype=torch.bfloat16); f=lambda: ops.gemm(a,w,o,bias=b,act=act)
    t=time_cuda(f,iters=20,warmup=3)
    print(f'  M={M} N={N} K={K} act={act} resid={res}: {t*1e3:7.4f} ms {2.0*M*N*K/t/1e12:7.1f} TF/s')
"
done
done 2>&1 | tee ../gpurun_out/ab_gemm_cluster.log
