run() { # name, env...
  name=$1; shift
  env "$@" timeout -k 10 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29600 bench.py --gpus 8 --mode train --steps 20 --warmup 5 2>gpurun_out/r2_train8_$name.err | grep '^{' > gpurun_out/r2_train8_$name.json
}
timeout -k 10 300 python bench.py --gpus 1 --mode train --steps 20 --warmup 5 2>/dev/null | grep '^{' > gpurun_out/r2_train1.json
run default HVLM_X=1
run plain HVLM_NCCL_MAX_CTAS=0 HVLM_NCCL_HIGH_PRIO=0
run ctas1 HVLM_NCCL_MAX_CTAS=1
run ctas8 HVLM_NCCL_MAX_CTAS=8
run pol1 HVLM_NCCL_MAX_CTAS=0 HVLM_NCCL_CTA_POLICY=1
python - <<'PY'
import json,glob
b=json.load(open('gpurun_out/r2_train1.json'))
print('N=1', b['value'], b['ms_per_step'])
for f in sorted(glob.glob('gpurun_out/r2_train8_*.json')):
    try:
        d=json.load(open(f)); a=d.get('allreduce',{})
        print(f.split('train8_')[1], d['value'], d['ms_per_step'], 'eff', round(d['value']/(8*b['value']),4), a.get('isolated_ms'), a.get('in_step_ms'))
    except Exception as e: print(f, 'ERR', e)
PY
