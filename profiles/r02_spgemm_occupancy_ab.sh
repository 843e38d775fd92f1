# A/B of the register caps of the numeric SpGEMM bins (run on the GPU box): FXII_LIB selects the variant library
# usage: bash profiles/r02_spgemm_occupancy_ab.sh base v3 v4 ...   (variants: build_variants/libfxii_<name>.so)
mkdir -p gpurun_out
for v in "$@"; do
  if [ $v = base ]; then unset FXII_LIB; else export FXII_LIB=$PWD/build_variants/libfxii_$v.so; fi
  python bench.py --steps 5 --no-cpu-baseline > gpurun_out/r2_sp_$v.json 2> gpurun_out/r2_sp_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_sp_$v.json')); print('$v', d['ms_per_step'], d['stage_ms']['spgemm_ms'], d['stage_ms']['spgemm_mp_ms'], d['e2e']['ms_per_step'])"
done
