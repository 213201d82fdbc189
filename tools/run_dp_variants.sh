# usage: bash tools/run_dp_variants.sh N   -- data-parallel schedule variants of the config-2 step on N GPUs (bench.py --quick)
N=${1:-2}
mkdir -p gpurun_out
run() { # tag, bench args, env...
  tag=$1; shift; bargs=$1; shift
  env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --quick --blocks 30 $bargs > gpurun_out/r02_dpv_n${N}_$tag.json 2> gpurun_out/r02_dpv_n${N}_$tag.err
  python -c "
import json,sys
try:
    p=json.load(open('gpurun_out/r02_dpv_n${N}_$tag.json'))
    print('$tag', round(p['ms_per_step']*1e3,1), 'us', p['dp_check']['ok'], p['dp'][40:110])
except Exception as e:
    print('$tag FAILED', e)"
}
run default_bucket_bf16_b16t512 "" AA_X=1
run single_bf16_b32t512 "--overlap 0" AA_DP_SINGLE=1 AA_AR_BLOCKS=32 AA_AR_THREADS=512
run bucket_bf16_b8t512 "" AA_AR_BLOCKS=8
run afterbptt_bf16_b16t512 "" AA_DP_SCHEDULE=after_bptt
