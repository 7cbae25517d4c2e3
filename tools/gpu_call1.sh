mkdir -p gpurun_out
(nproc; python -c "import os;print('affinity',len(os.sched_getaffinity(0)),'cpu_count',os.cpu_count())"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /proc/loadavg; nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv) > gpurun_out/hostinfo.txt 2>&1
bash tools/gpu_next.sh
