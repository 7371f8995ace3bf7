set -x
python scripts/profile_case.py > gpurun_out/plain_prof.log 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 5 --launch-count 5 -k regex:"nms_qc|osd" -o gpurun_out/prof_final -f python scripts/profile_case.py > gpurun_out/ncu_final.log 2>&1
python scripts/blockmin_perf.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:osd_blocks -o gpurun_out/prof_final_blocks -f python scripts/blockmin_perf.py > gpurun_out/ncu_final_blocks.log 2>&1
python scripts/profile_pb.py 2 65536 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 1 --launch-count 1 -k regex:osd_pb -o gpurun_out/prof_final_pb -f python scripts/profile_pb.py 2 65536 > gpurun_out/ncu_final_pb.log 2>&1
python scripts/c1_dropin.py > gpurun_out/f_c1_dropin.json 2> gpurun_out/f_c1_dropin.err
python scripts/dl_perf.py > gpurun_out/f_dl.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
