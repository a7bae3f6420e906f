mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv3_tc_kdn_kernel|conv3_tc_kernel|conv3_wgrad_dsh" -s 5 -c 5 -o gpurun_out/r2_fullres_final python tools/fullres_prof.py > gpurun_out/r2_fullres_final.log 2>&1
ncu -i gpurun_out/r2_fullres_final.ncu-rep --page raw --csv > gpurun_out/r2_fullres_final_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2_fullres_final_raw.csv')))
hdr=rows[0]; units=rows[1]; data=rows[2:]
want=["Kernel Name","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","l1tex__m_xbar2l1tex_read_bytes.sum","launch__registers_per_thread","smsp__cycles_active.avg"]
for wn in want:
    if wn in hdr:
        i=hdr.index(wn)
        print("%-80s %s   [%s]"%(wn, " | ".join(r[i][:34] for r in data), units[i]))
PY
