run() { tag=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err; python -c "
import json; d=json.load(open('gpurun_out/ab_$tag.json')); print('$tag', round(d['value'],3), round(d['e2e']['value'],3), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"; }
run fuse1 LLAVAMOD_FUSE_RESIDUAL=1
run fuse0 LLAVAMOD_FUSE_RESIDUAL=0
run fuse0_nocap LLAVAMOD_FUSE_RESIDUAL=0 LMOD_ATTN_REGCAP=0
run fuse0_nokeep LLAVAMOD_FUSE_RESIDUAL=0 LMOD_KL_KEEP=0
run fuse1_b LLAVAMOD_FUSE_RESIDUAL=1
run fuse0_b LLAVAMOD_FUSE_RESIDUAL=0
