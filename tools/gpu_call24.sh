#!/bin/bash
# Round 2, N-GPU call (one box, N = $1): weak-scaling batch bench, chunk streaming (configs[4]: 512 channels dealt over the
# ranks), configs[2]/[3] from multi-read fast5 files through MapPool, and the CLI itself under torchrun on a fast5 directory.
N=${1:-4}
BF=${2:-$((16 * N))}      # fast5 files of the bench pass
CF=${3:-$((16 * N))}      # fast5 files in the CLI's directory
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --no-extras > gpurun_out/bench24_${N}gpu.json 2> gpurun_out/bench24_${N}gpu.err; echo "batch rc=$?"; cut -c1-900 gpurun_out/bench24_${N}gpu.json; tail -2 gpurun_out/bench24_${N}gpu.err
timeout 600 $TR bench.py --gpus $N --workload stream --steps 2 --warmup 1 > gpurun_out/bench24_stream_${N}gpu.json 2> gpurun_out/bench24_stream_${N}gpu.err; echo "stream rc=$?"; cut -c1-1300 gpurun_out/bench24_stream_${N}gpu.json; tail -2 gpurun_out/bench24_stream_${N}gpu.err
timeout 900 $TR bench.py --gpus $N --workload fast5 --files $BF --steps 1 --warmup 1 > gpurun_out/bench24_fast5_${N}gpu.json 2> gpurun_out/bench24_fast5_${N}gpu.err; echo "fast5 rc=$?"; cut -c1-1300 gpurun_out/bench24_fast5_${N}gpu.json; tail -2 gpurun_out/bench24_fast5_${N}gpu.err
# the CLI: `uncalled map` on a directory of fast5 files, one process per GPU
mkdir -p /tmp/f5dir; for i in $(seq 1 $CF); do ln -sf "$PWD/bench_data/bench_reads_4000x4000.fast5" /tmp/f5dir/batch_$i.fast5; done
T0=$(date +%s.%N)
timeout 600 $TR -m uncalled_b200 map bench_data/g4m7 /tmp/f5dir > /tmp/cli_out.paf 2> gpurun_out/cli24_${N}gpu.err; echo "cli rc=$?"
T1=$(date +%s.%N)
python - <<PY
n = sum(1 for l in open("/tmp/cli_out.paf") if l.strip())
m = sum(1 for l in open("/tmp/cli_out.paf") if l.strip() and l.split("\t")[5] != "*")
dt = $T1 - $T0
import json
line = {"what": "python -m torch.distributed.run --nproc-per-node $N -m uncalled_b200 map bench_data/g4m7 <dir of %d multi-read fast5 files>" % $CF,
        "n_gpus": $N, "paf_lines": n, "mapped": m, "wall_s_incl_startup": dt, "reads_per_s_incl_startup": n / dt}
print(json.dumps(line)); open("gpurun_out/cli24_${N}gpu.json", "w").write(json.dumps(line) + "\n")
PY
tail -3 gpurun_out/cli24_${N}gpu.err
