# round-2 session on 8 B200 (gpurun --gpus 8): world-8 parity at HEAD, the driver's bench configuration three times, 200 steps, GMRES(50)
set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_gpu_dist.py -q -k "test_row_partitioned and 8- and not helpers" 2>&1 | tail -3
for i in 3 4 5; do timeout 300 $TR --master-port 2951$i bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_n8_steps20_run$i.json 2> gpurun_out/r2_n8_steps20_run$i.err; echo "rc=$?"; grep "^{" gpurun_out/r2_n8_steps20_run$i.json | head -c 250; echo; done
timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 10 --no-e2e > gpurun_out/r2_n8_steps200_final.json 2> gpurun_out/r2_n8_steps200_final.err; echo "rc=$?"; grep "^{" gpurun_out/r2_n8_steps200_final.json | head -c 250; echo
for p in 1 0; do HB_GMRES_PIPELINE=$p timeout 300 $TR --master-port 2953$p bench.py --gpus 8 --workload gmres --steps 250 > gpurun_out/r2_n8_gmres_pipeline$p.json 2> gpurun_out/r2_n8_gmres_pipeline$p.err; echo "rc=$?"; grep "^{" gpurun_out/r2_n8_gmres_pipeline$p.json | head -c 250; echo; done
