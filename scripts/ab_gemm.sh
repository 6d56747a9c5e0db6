#!/bin/bash
# A/B GEMM variants built under pi3_slam_b200/build/variants on the same box
mkdir -p gpurun_out
cd scripts
for rep in 1 2; do
for v in ../pi3_slam_b200/build/variants/*.so; do
  echo "== $v (rep $rep)"
  PI3B_LIB=$(realpath $v) timeout -s KILL 100 python -c "
import torch, sys
sys.path.insert(0,'..')
from pi3_slam_b200 import ops
from bench_kernels import time_cuda
dev='cuda'
for (M,N,K,act,res) in [(100400,3072,1024,0,0),(100400,1024,1024,0,1),(100400,4096,1024,1,0),(100400,1024,4096,0,1)]:
    a=torch.randn(M,K,device=dev).bfloat16(); w=(torch.randn(N,K,device=dev)*0.02).bfloat16(); b=torch.randn(N,device=dev); g=torch.rand(N,device=dev)
    if res:
        x=torch.randn(M,N,device=dev); f=lambda: ops.gemm(a,w,x,bias=b,gamma=g,resid=x)
    else:
        o=torch.empty(M,N,device=dev,dtype=torch.bfloat16); f=lambda: ops.gemm(a,w,o,bias=b,act=act)
    t=time_cuda(f,iters=20,warmup=3)
    print(f'  M={M} N={N} K={K} act={act} resid={res}: {t*1e3:7.4f} ms {2.0*M*N*K/t/1e12:7.1f} TF/s')
# the decoder's fused QKV projection (q/k LayerNorm + RoPE in the epilogue)
M,T,pw=100400,1004,37
qkn=torch.cat([torch.rand(64,device=dev)+0.5,torch.randn(64,device=dev)*0.1,torch.rand(64,device=dev)+0.5,torch.randn(64,device=dev)*0.1]).contiguous()
inv=1.0/(100.0**(torch.arange(0,32,2).float()/32)); ang=torch.einsum('i,j->ij',torch.arange(pw+2,dtype=torch.float32),inv)
cs=torch.stack((ang.cos(),ang.sin()),dim=-1).contiguous().to(dev)
a=torch.randn(M,1024,device=dev).bfloat16(); w=(torch.randn(3072,1024,device=dev)*0.02).bfloat16(); b=torch.randn(3072,device=dev)
o=torch.empty(M,3072,device=dev,dtype=torch.bfloat16)
t=time_cuda(lambda: ops.gemm_qkv(a,w,o,b,16,qkn,1e-5,cs,T,5,pw),iters=20,warmup=3)
print(f'  fused qkv + q/k LayerNorm + RoPE: {t*1e3:7.4f} ms {2.0*M*3072*1024/t/1e12:7.1f} TF/s')
"
done
done 2>&1 | tee ../gpurun_out/ab_gemm.log
