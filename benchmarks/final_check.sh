# Round-end verification on one B200: tests, bench (both arms), sweeps, projection kernel, RoBERTa, ncu.
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 | tail -1 | cut -c1-300
python benchmarks/sweep_functions.py --bits 3 --md gpurun_out/sweep3.md --json gpurun_out/sweep3.json > /dev/null
python benchmarks/sweep_functions.py --bits 1,2,4,5,6,7,8 --md gpurun_out/sweep_rest.md --json gpurun_out/sweep_rest.json > /dev/null
python benchmarks/sketch_bench.py > gpurun_out/sketch_bench.json
python benchmarks/roberta_step.py --dtype fp32 > gpurun_out/roberta_fp32.txt 2>&1; cat gpurun_out/roberta_fp32.txt
python benchmarks/roberta_step.py --dtype bf16 > gpurun_out/roberta_bf16.txt 2>&1; cat gpurun_out/roberta_bf16.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"tiles_kernel|sketch_kernel" -c 14 -o gpurun_out/prof_final python benchmarks/profile_kernels.py 1 > /dev/null 2>&1
ls -la gpurun_out
