set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_dist.py -q -k "test_row_partitioned and 8- and not helpers" 2>&1 | tail -4
for i in 1 2; do timeout 300 $TR --master-port 2951$i bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_n8_steps20_run$i.json 2> gpurun_out/r2_n8_steps20_run$i.err; echo "rc=$?"; tail -c 1200 gpurun_out/r2_n8_steps20_run$i.json | head -c 1200; done
for d in 1 0; do HB_PEER_HALO_DEFER=$d timeout 300 $TR --master-port 2952$d bench.py --gpus 8 --steps 200 --warmup 10 --no-e2e > gpurun_out/r2_n8_steps200_defer$d.json 2> gpurun_out/r2_n8_steps200_defer$d.err; echo "rc=$?"; head -c 300 gpurun_out/r2_n8_steps200_defer$d.json; done
HB_DIST_PEER=0 timeout 300 $TR --master-port 29530 bench.py --gpus 8 --steps 200 --warmup 10 --no-e2e > gpurun_out/r2_n8_steps200_nccl.json 2> gpurun_out/r2_n8_steps200_nccl.err; echo "rc=$?"; head -c 300 gpurun_out/r2_n8_steps200_nccl.json
timeout 300 $TR --master-port 29531 bench.py --gpus 8 --workload gmres --steps 250 > gpurun_out/r2_n8_gmres.json 2> gpurun_out/r2_n8_gmres.err; echo "rc=$?"; head -c 400 gpurun_out/r2_n8_gmres.json
