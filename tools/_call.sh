TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/mp_parity.py > gpurun_out/r2b_parity2.log 2>&1; echo "parity rc $?" >> gpurun_out/r2b_parity2.log
timeout 300 $TR --master-port 29512 tools/mp_profile.py 1 > gpurun_out/r2b_prof2.log 2>&1
FLIGHT_DEPTHS=1,2,4,8 FLIGHT_ENVS="GXY_GEN_SKIP_FAR=1;GXY_GEN_SKIP_FAR=0;GXY_PEER_OVERLAP=0" timeout 600 $TR --master-port 29513 tools/flight_sweep.py 1 24 > gpurun_out/r2b_sweep2.log 2>&1
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2b_bench2.json 2> gpurun_out/r2b_bench2.err
