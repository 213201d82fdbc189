mkdir -p gpurun_out
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --quick --blocks 30 > gpurun_out/r02_n8_$tag.json 2> gpurun_out/r02_n8_$tag.err
  echo "$tag rc $?"; python -c "
import json,sys
p=json.load(open('gpurun_out/r02_n8_$tag.json'))
print('$tag', round(p['ms_per_step']*1e3,1), 'us', p['dp_check']['ok'])"
}
run b8t256 AA_AR_BLOCKS=8 AA_AR_THREADS=256
run b8t512 AA_AR_BLOCKS=8 AA_AR_THREADS=512
run b4t256 AA_AR_BLOCKS=4 AA_AR_THREADS=256
run b16t128 AA_AR_BLOCKS=16 AA_AR_THREADS=128
run b32t512 AA_AR_BLOCKS=32 AA_AR_THREADS=512
AA_AR_BLOCKS=8 AA_AR_THREADS=256 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/timeline_train.py > gpurun_out/r02_timeline_n8_v2.txt 2> /dev/null; echo tl rc $?
