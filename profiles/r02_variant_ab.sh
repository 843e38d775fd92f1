# A/B of compile-time variants of the library (run on the GPU box): FXII_LIB selects build_variants/libfxii_<name>.so
# usage: bash profiles/r02_variant_ab.sh base pf0 ...
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v = base ]; then unset FXII_LIB; else export FXII_LIB=$PWD/build_variants/libfxii_$v.so; fi
  python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r2_ab_$v.json 2> gpurun_out/r2_ab_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_ab_$v.json')); s=d['stage_ms']; print('$v', 'step', round(d['ms_per_step'],3), 'rows', round(s['ms_locate'],3), 'pi', round(s['pi_ms'],3), 'transpose', round(s['transpose_ms'],3), 'spgemm', round(s['spgemm_ms'],3), 'mpi', round(s['spgemm_mp_ms'],3), 'e2e', round(d['e2e']['ms_per_step'],2))"
done
