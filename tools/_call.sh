TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
FLIGHT_DEPTHS=8,16 FLIGHT_ENVS="GXY_RAYS_PER_THREAD=0;GXY_RAYS_PER_THREAD=2;GXY_RAYS_PER_THREAD=4;GXY_RAYS_PER_THREAD=8;GXY_RAYS_PER_THREAD=16" timeout 600 $TR --master-port 29513 tools/flight_sweep.py 1 32 > gpurun_out/r2l_sweep4.log 2>&1
timeout 600 $TR --master-port 29511 tools/mp_parity.py > gpurun_out/r2l_parity4.log 2>&1; echo "parity rc $?" >> gpurun_out/r2l_parity4.log
