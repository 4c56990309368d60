"""Turn an `ncu --set full` report of the DCN gather-GEMM launches of one frame into profiles/r2_dcn_ncu_traffic.json
(run where ncu is installed: python tools/ncu_traffic.py <report.ncu-rep> <out.json>)."""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, name, scale_from_unit=True):
    v = r[col[name]].replace(",", "")
    x = float(v) if v not in ("", "n/a") else 0.0
    u = units[col[name]]
    if scale_from_unit:
        x *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3}.get(u, 1.0)
    return x
layers = []
for i, r in enumerate(rows[2:]):
    layers.append({"launch": i, "kernel": r[col["Kernel Name"]][:60], "grid": r[col["Grid Size"]], "us_under_ncu": val(r, "gpu__time_duration.sum"),
                   "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                   "tensor_pipe_active_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False),
                   "l1_hit_pct": val(r, "l1tex__t_sector_hit_rate.pct", False),
                   "smem_config_kb": val(r, "launch__shared_mem_config_size", False)})
tot = sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in layers)
json.dump({"source": "ncu --set full --clock-control none -k regex:TsDcnAP -c 16 python bench.py --steps 2 --warmup 3 --no-graph --depth 1 --no-cpu-baseline (round 2, final build)",
           "kernel": "tc_ts_kernel<TsDcnAP,64|128,6> (the gather-GEMM launch of deft_dcnv2_forward, 3xFP16), the 16 DCN layers of one frame",
           "note": "dram__bytes_write is ~0 under ncu's per-kernel replay: outputs stay in the 126 MB L2; reads are the cold-cache input + weights",
           "dram_bytes_per_frame": tot, "dram_bytes_per_launch_avg": tot / max(len(layers), 1), "layers": layers}, open(out, "w"), indent=1)
print("wrote", out, len(layers), "launches", tot / 1e6, "MB")
