# Round-2 final single-GPU measurements (run on the GPU box): bench line, reference arm, ncu launch list, ncu captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r02_final_clocks.csv &
SMI=$!
python bench.py > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err; echo "reference rc=$?"
kill $SMI
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_ncu_list.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:spgemm_warp -s 66 -c 12 -o gpurun_out/r02_spgemm_final \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_ncu_sp.log 2>&1; echo "ncu spgemm rc=$?"
ncu --set full --clock-control none --import-source on -k regex:pi_fused -s 4 -c 1 -o gpurun_out/r02_pi_fused_final \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r02_final_ncu_pi.log 2>&1; echo "ncu pi rc=$?"
tail -c 600 gpurun_out/r02_final_ref.json
