cd /root/repo; mkdir -p gpurun_out
python tools/pre_once.py 256 > gpurun_out/r02_pre_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:resample_fused -s 3 -c 1 -f -o gpurun_out/r02_pre_fused python tools/pre_once.py 256 > gpurun_out/r02_pre_ncu2.log 2>&1
cat gpurun_out/r02_pre_plain.log
