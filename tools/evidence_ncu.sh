#!/bin/bash
# round-2 evidence: launch list of the default bench, ncu --set full of the step's two kernels on the C2 batch and on the C5 per-GPU share,
# of the C3 chain, and of the varint kernels.  Reports are exported to CSV on the box (gpurun brings back 64 MiB at most); only the
# two-launch C5-share report travels as .ncu-rep.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cat > gpurun_out/how_r02.json <<'J'
{"launch_list": "ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kernel' -c 400 --csv python bench.py --steps 4 --warmup 3 --no-cpu  (eager warm-up steps + graph replays of the timed region + the short c3/c4/c5 passes + the single-request pass)",
 "full_capture": "ncu --set full --clock-control none --import-source on -k regex:<kernels> -s <skip> -c <n> python bench.py --workload <w> [--batch 1024] --steps 2 --warmup 3 --no-cpu --no-extra ; varint: python tools/varint_probe.py --only mixed --reps 1 ; exported with ncu -i <rep> --page raw --csv"}
J
export_rep() {  # $1 = name (without extension), $2 = keep the report? (1/0), $3 = kernel whose source page to export (optional)
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1.raw.csv 2>/dev/null
  if [ -n "$3" ]; then ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source cuda --kernel-name "$3" --launch-count 1 > gpurun_out/$1.source.csv 2>/dev/null; fi
  if [ "$2" != "1" ]; then rm -f gpurun_out/$1.ncu-rep; fi
}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kernel' -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 4 --warmup 3 --no-cpu > gpurun_out/rp_launches.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^move_kernel$|decode_fused_staged_kernel' -s 2 -c 2 -f -o gpurun_out/prof_r02_c2 python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/rp_c2.log 2>&1; echo "c2 rc=$?"
export_rep prof_r02_c2 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'^move_kernel$|decode_fused_staged_kernel' -s 2 -c 2 -f -o gpurun_out/prof_r02_c5share python bench.py --workload c5 --batch 1024 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/rp_c5.log 2>&1; echo "c5share rc=$?"
export_rep prof_r02_c5share 1
timeout 900 ncu --set full --clock-control none -k regex:'venc_len|frame_requests|^move_kernel$|venc_emit|venc_fused|decode_fused_kernel' -s 10 -c 8 -f -o gpurun_out/prof_r02_c3 python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/rp_c3.log 2>&1; echo "c3 rc=$?"
export_rep prof_r02_c3 0
timeout 900 ncu --set full --clock-control none -k regex:'decode_fused_cast|move_guarded|^move_kernel$' -s 6 -c 4 -f -o gpurun_out/prof_r02_c4 python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/rp_c4.log 2>&1; echo "c4 rc=$?"
export_rep prof_r02_c4 0
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'venc|vdec' -s 8 -c 8 -f -o gpurun_out/prof_r02_varint python tools/varint_probe.py --only mixed --reps 1 > gpurun_out/rp_varint.log 2>&1; echo "varint rc=$?"
export_rep prof_r02_varint 0
du -sh gpurun_out; ls -la gpurun_out | head -60
