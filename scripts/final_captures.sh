# The exact command list of the round's final GPU run (under gpurun, one B200).  Everything lands in gpurun_out/;
# scripts/refresh_profiles.sh (run on the CPU box afterwards) turns it into the files under profiles/.
set -x
python -m pytest tests -x -q -m gpu > gpurun_out/f_pytest.log 2>&1; tail -3 gpurun_out/f_pytest.log
# ncu capture of the four decoder kernels; the stamped traffic file is made right here so that the bench line below reads
# the per-frame counts of the very build it times
python scripts/profile_case.py > gpurun_out/plain_prof.log 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 5 --launch-count 5 -k regex:"nms_qc|osd" -o gpurun_out/prof_final -f python scripts/profile_case.py > gpurun_out/ncu_final.log 2>&1
python scripts/make_traffic.py gpurun_out/prof_final.ncu-rep gpurun_out/prof_stamp.json > /dev/null && cp profiles/r02_traffic.json gpurun_out/f_traffic.json
python bench.py --impl reference > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
python bench.py --steps 2 --warmup 3 --frames 262144 --no-configs --no-cpu-baseline > gpurun_out/f_b262.json 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 2 --warmup 3 --frames 262144 --no-configs --no-cpu-baseline > gpurun_out/f_ncu_b262.log 2>&1
for o in 1 2 3; do python -m short_ldpc_decoding_osd_b200.simulate --frames 100000000 --order $o > gpurun_out/f_fer_osd$o.jsonl 2> gpurun_out/f_fer_osd$o.err; done
python scripts/blockmin_perf.py > gpurun_out/f_bm.log 2>&1
python scripts/batch_sweep.py > gpurun_out/f_batch_sweep.jsonl 2> gpurun_out/f_batch_sweep.err
python scripts/blockmin_perf.py > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 7 --launch-count 1 -k regex:osd_blocks -o gpurun_out/prof_final_blocks -f python scripts/blockmin_perf.py > gpurun_out/ncu_final_blocks.log 2>&1
python scripts/profile_pb.py 2 65536 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on --launch-skip 1 --launch-count 1 -k regex:osd_pb -o gpurun_out/prof_final_pb -f python scripts/profile_pb.py 2 65536 > gpurun_out/ncu_final_pb.log 2>&1
python scripts/c1_dropin.py > gpurun_out/f_c1_dropin.json 2> gpurun_out/f_c1_dropin.err
python scripts/dl_perf.py > gpurun_out/f_dl.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -5
