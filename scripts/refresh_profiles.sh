#!/bin/bash
# Run HERE (CPU box) after scripts/final_captures.sh has come back through gpurun_out/: turns the captures into the
# text / JSON files under profiles/.
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/f_traffic.json profiles/r02_traffic.json  # made on the GPU box right after the capture (scripts/final_captures.sh)
python scripts/ncu_kernel_report.py gpurun_out/prof_final.ncu-rep nms_qc 131072 > profiles/r02_ncu_nms_qc.txt
python scripts/ncu_kernel_report.py gpurun_out/prof_final.ncu-rep osd_pair 131072 > profiles/r02_ncu_osd_pair.txt
python scripts/ncu_kernel_report.py gpurun_out/prof_final.ncu-rep osd_kernel 131072 > profiles/r02_ncu_osd_order1.txt
python scripts/ncu_kernel_report.py gpurun_out/prof_final.ncu-rep osd3 16384 > profiles/r02_ncu_osd3.txt
python scripts/ncu_kernel_report.py gpurun_out/prof_final_blocks.ncu-rep osd_blocks 244859 > profiles/r02_ncu_osd_blocks.txt
python scripts/ncu_kernel_report.py gpurun_out/prof_final_pb.ncu-rep osd_pb 65536 > profiles/r02_ncu_osd_pb_order2.txt
python scripts/sass_opcodes.py > profiles/r02_sass_opcodes.txt
cp gpurun_out/f_bench.json profiles/r02_bench_1gpu.json
cp gpurun_out/f_bench_ref.json profiles/r02_bench_reference_arm_1gpu.json
cp gpurun_out/f_batch_sweep.jsonl profiles/r02_batch_sweep_1gpu.jsonl
cp gpurun_out/f_launches.csv profiles/r02_launches_bench_262144.csv
cp gpurun_out/f_c1_dropin.json profiles/r02_config1_dropin_api.json
cp gpurun_out/f_dl.log profiles/r02_dl_scheme_perf.txt
cp gpurun_out/f_bm.log profiles/r02_block_minima_perf.txt
for o in 1 2 3; do cp gpurun_out/f_fer_osd$o.jsonl profiles/r02_fer_nms_osd${o}_1e8_frames.jsonl; done
echo refreshed
