#!/bin/bash
# A/B the attention kernel variants built under pi3_slam_b200/build/variants (same process conditions, same box)
mkdir -p gpurun_out
cd scripts
for rep in ${REPS:-1 2}; do
for v in ../pi3_slam_b200/build/variants/*.so; do
  echo "== $v (rep $rep)"
  PI3B_LIB=$(realpath $v) timeout -s KILL 100 python -c "
import torch, sys
sys.path.insert(0,'..')
from pi3_slam_b200 import ops
from bench_kernels import time_cuda
for (B,S,H,it) in [(100,1004,16,10),(1,64300,16,4),(1,100400,16,4)]:
    qkv=torch.randn(B*S,3*H*64,device='cuda').bfloat16(); out=torch.empty(B*S,H*64,device='cuda',dtype=torch.bfloat16)
    t=time_cuda(lambda: ops.attention_qkv(qkv,out,B,S,H),iters=it,warmup=2)
    print(f'  B={B} S={S}: {t*1e3:8.3f} ms  {4.0*B*H*S*S*64/t/1e12:7.1f} TF/s')
"
done
done 2>&1 | tee ../gpurun_out/ab_attn.log
