#!/bin/bash
# 2-GPU sanity of the torchrun path (+ an ncu --set full capture at the real batch size on GPU 0)
TAG=${1:-r01m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $OUT/gpus.txt
echo "== ncu full capture at batch 1024"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k1_idct8_tma|k2_ycbcr420" -s 8 -c 2 -o $OUT/prof1024 \
    python bench.py --steps 2 --warmup 3 --e2e-batch 16 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
python scripts/ncu_traffic.py $OUT/prof1024.ncu-rep cfg2 1024 | tee $OUT/traffic.json
cp profiles/traffic.json $OUT/traffic_file.json
echo "== N=1"; timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --cpu-seconds 3 2>$OUT/err1.txt | tee $OUT/bench_n1.json
echo "== N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 2>$OUT/err2.txt | tee $OUT/bench_n2.json
tail -3 $OUT/err2.txt
echo "== N=2 reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 0 2>>$OUT/err2.txt | tee $OUT/bench_ref_n2.json
ls -la $OUT
