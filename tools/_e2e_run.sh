run() {
  tag=$1; shift
  env "$@" python bench.py --steps 3 --warmup 2 --extras none --no-cpu-baseline > gpurun_out/e2e_$tag.json 2> gpurun_out/e2e_$tag.err
  python - <<P
import json
d=json.loads(open('gpurun_out/e2e_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value'],2), round(d['e2e']['value'],2), round(d['e2e']['pack_ms'],1), round(d['e2e']['unpack_ms'],1))
P
}
run w2r1 ZG_UNPACK_WORKERS=2
run w2r0 ZG_UNPACK_WORKERS=2 ZG_UNPACK_RAMP=0
run w1r1 ZG_UNPACK_WORKERS=1
run w3r1 ZG_UNPACK_WORKERS=3
run p512 ZG_PACK_SLICE_MB=512
run p1024 ZG_PACK_SLICE_MB=1024
