#!/bin/bash
# round 2: N-GPU runs (bench.py weak scaling; key switch with fused exchanges and with NCCL all-gathers)
NG=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29655"
$TR bench.py --gpus $NG --steps 10 --warmup 3 > gpurun_out/r02_bench_${NG}gpu.json 2> gpurun_out/r02_bench_${NG}gpu.err
{
echo "# tools/bench_keyswitch.py on $NG B200 (torchrun, one process per GPU), round 2, final kernels"
echo "## --fused --verify --profile"
$TR tools/bench_keyswitch.py --fused --verify --profile 2>&1 | grep -E "^\{|^rank"
echo "## --verify (NCCL all-gathers)"
$TR tools/bench_keyswitch.py --verify 2>&1 | grep -E "^\{"
} > gpurun_out/r02_keyswitch_${NG}gpu.txt 2>&1
cat gpurun_out/r02_bench_${NG}gpu.json | cut -c1-330; cat gpurun_out/r02_keyswitch_${NG}gpu.txt | cut -c1-260
