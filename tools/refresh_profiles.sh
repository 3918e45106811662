# Copy the outputs of benchmarks/final_check.sh (gpurun_out/) into the tracked profiles/ files.
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/bench_n1.json profiles/r01_bench_n1.json
cp gpurun_out/sweep3.md profiles/r01_function_sweep_3bit.md; cp gpurun_out/sweep3.json profiles/r01_function_sweep_3bit.json
cp gpurun_out/sweep_rest.md profiles/r01_function_sweep_other_bits.md; cp gpurun_out/sweep_rest.json profiles/r01_function_sweep_other_bits.json
cp gpurun_out/sketch_bench.json profiles/r01_sketch_bench.json
cp gpurun_out/roberta_fp32.txt profiles/r01_roberta_fp32.txt; cp gpurun_out/roberta_bf16.txt profiles/r01_roberta_bf16.txt
cp gpurun_out/launches.csv profiles/r01_launches_bench.csv
mkdir -p build/scratch
ncu -i gpurun_out/prof_final.ncu-rep --page raw --csv > build/scratch/prof_final.csv 2>/dev/null
python - <<'PY'
import subprocess
hdr = '''# ncu launch list of the bench command (round 1)

`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1` on a B200 (first 400 launches: input generation, warm-up + 2 timed steps, then the first chunks of the host-staged e2e pass; summary by `tools/launch_list.py`). Per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes.

'''
body = subprocess.run(['python', 'tools/launch_list.py', 'gpurun_out/launches.csv'], capture_output=True, text=True).stdout
open('profiles/r01_launches_bench.md', 'w').write(hdr + body)
tab = subprocess.run(['python', 'tools/ncu_table.py', 'build/scratch/prof_final.csv', '--traffic', 'profiles/ncu_traffic.json'],
                     capture_output=True, text=True).stdout
old = open('profiles/r01_ncu_final_kernels.md').read()
head = old[:old.index('| kernel |')]
rest = old[old.index('| kernel |'):]
lines = rest.split('\n')
i = 0
while i < len(lines) and lines[i].startswith('|'):
    i += 1
open('profiles/r01_ncu_final_kernels.md', 'w').write(head + tab + '\n'.join(lines[i:]))
PY
