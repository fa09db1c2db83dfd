#!/bin/bash
# Turn the raw outputs of profiles/make_r01.sh (gpurun_out/) into the committed summaries.
set -e
python profiles/ncu_summary.py gpurun_out/r01_step_kernel.ncu-rep > profiles/r01_step_kernel_ncu_summary.txt
python - <<'PY'
import json,csv,subprocess
out=subprocess.run(["ncu","-i","gpurun_out/r01_step_kernel.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); hdr=rows[0]; units=rows[1]
def val(r,k):
    i=hdr.index(k); v=float(r[i]); u=units[i]
    return v*{'Mbyte':1e6,'Kbyte':1e3,'byte':1,'Gbyte':1e9}.get(u,1)
launches=[dict(us=float(r[hdr.index('gpu__time_duration.sum')]), dram_read=val(r,'dram__bytes_read.sum'), dram_write=val(r,'dram__bytes_write.sum'), inst=float(r[hdr.index('smsp__inst_executed.sum')])) for r in rows[2:]]
mean=sum(l['dram_read']+l['dram_write'] for l in launches)/len(launches)
json.dump(dict(source="ncu --set full --clock-control none, profiles/make_r01.sh, report gpurun_out/r01_step_kernel.ncu-rep (launch 8 = defender turn, launch 9 = attacker turn)",
       envs_per_launch=65536, device_slots=100, dram_bytes_per_launch=mean, launches=launches,
       note="dram write bytes of a single captured launch are small because the 30.7 MB bulk store stays in the 126 MB L2 until evicted; algorithmic bytes per launch = 67.4 MB, record traffic = 2 x 30.7 MB"),open('profiles/r01_step_kernel_traffic.json','w'),indent=1)
PY
python - <<'PY'
import json,csv,subprocess
out=subprocess.run(["ncu","-i","gpurun_out/r01_step_kernel_fused.ncu-rep","--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); hdr=rows[0]; units=rows[1]
def val(r,k):
    i=hdr.index(k); v=float(r[i]); u=units[i]
    return v*{'Mbyte':1e6,'Kbyte':1e3,'byte':1,'Gbyte':1e9}.get(u,1)
launches=[dict(us=float(r[hdr.index('gpu__time_duration.sum')]), dram_read=val(r,'dram__bytes_read.sum'), dram_write=val(r,'dram__bytes_write.sum'), inst=float(r[hdr.index('smsp__inst_executed.sum')])) for r in rows[2:]]
mean=sum(l['dram_read']+l['dram_write'] for l in launches)/len(launches)
json.dump(dict(source="ncu --set full --clock-control none, profiles/make_r01.sh, report gpurun_out/r01_step_kernel_fused.ncu-rep (two launches of 4 fused plain steps each)",
       envs_per_launch=65536, device_slots=100, steps_per_launch=4, dram_bytes_per_launch=mean, launches=launches,
       note="4 steps per launch: the 30.7 MB record span is read once and written once per launch (writes of one captured launch mostly stay in L2), plus 4 x 2 MB of action rows; algorithmic bytes per launch = 4 x 67.4 MB"),open('profiles/r01_step_kernel_fused_traffic.json','w'),indent=1)
PY
python profiles/ncu_summary.py gpurun_out/r01_step_kernel_fused.ncu-rep > profiles/r01_step_kernel_fused_ncu_summary.txt
for k in 1 2; do ncu -i gpurun_out/r01_step_kernel.ncu-rep --page source --csv --print-source cuda,sass --kernel-id :::$k 2>/dev/null > gpurun_out/src_r01_$k.csv; done
python profiles/ncu_lines.py gpurun_out/src_r01_1.csv 25 > profiles/r01_step_kernel_defender_lines.txt 2>/dev/null || true
python profiles/ncu_lines.py gpurun_out/src_r01_2.csv 25 > profiles/r01_step_kernel_attacker_lines.txt 2>/dev/null || true
cp gpurun_out/r01_bench.json profiles/r01_bench_line.json
cp gpurun_out/r01_launches.csv profiles/r01_launches.csv
