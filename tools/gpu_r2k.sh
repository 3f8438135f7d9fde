set -x
cd /root/repo
mkdir -p gpurun_out/r2k
timeout 300 python tools/tc_timeline.py --H 256 > gpurun_out/r2k/timeline.txt 2>&1
timeout 300 python tools/tc_timeline.py --H 128 >> gpurun_out/r2k/timeline.txt 2>&1
cat gpurun_out/r2k/timeline.txt
