#!/bin/bash
# BASELINE.json configs[2] / configs[3] and the VQA line on N GPUs of one box (gpurun --gpus N):  N=${N:-2}
cd "$(dirname "$0")/../.."
N=${N:-2}
CAP=${CAP:-1000}
IMG=${IMG:-1000}
mkdir -p gpurun_out/r2
O=gpurun_out/r2
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/s5_env_n$N.txt
timeout 600 $RUN bench.py --gpus $N --workload multitask --steps 100 --warmup 5 > $O/s5_multitask_n$N.json 2> $O/s5_multitask_n$N.err
rm -f $O/s5_nccl_raw_n$N.*; NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT NCCL_DEBUG_FILE=$O/s5_nccl_raw_n$N.%p timeout 900 $RUN bench.py --gpus $N --workload retrieval --captions $CAP --images $IMG --steps 1 --warmup 1 --batch 256 \
    > $O/s5_retrieval_n$N.json 2> $O/s5_retrieval_n$N.err
cat $O/s5_nccl_raw_n$N.* 2>/dev/null | grep -E "nranks|NVLS|Init COMPLETE|NCCL version" | sort | uniq -c | sort -rn | head -60 | cut -c1-260 > $O/s5_nccl_n$N.txt; rm -f $O/s5_nccl_raw_n$N.*
timeout 600 $RUN bench.py --gpus $N --steps 100 --warmup 5 > $O/s5_vqa_n$N.json 2> $O/s5_vqa_n$N.err
for w in multitask retrieval vqa; do echo "== $w"; cut -c1-1500 $O/s5_${w}_n$N.json; tail -n 4 $O/s5_${w}_n$N.err | cut -c1-300; done
wc -l $O/s5_nccl_n$N.txt; head -5 $O/s5_nccl_n$N.txt | cut -c1-200
