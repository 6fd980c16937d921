"""`ncu --set full --csv --page raw` log -> one compact table row per launch (the per-round summaries under profiles/).
usage: python tools/ncu_table.py gpurun_out/x.csv [kernel-substring] > profiles/rNN_ncu_x.txt"""
import csv, re, sys

path = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
lines = [l for l in open(path) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr, units, body = rows[0], rows[1], rows[2:]
col = lambda p: next((i for i, h in enumerate(hdr) if re.search(p, h)), None)
C = dict(name=col(r"^Kernel Name$"), grid=col(r"^Grid Size$"), t=col(r"gpu__time_duration\.sum$"), tensor=col(r"sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct"),
         tensor2=col(r"sm__inst_executed_pipe_tensor.*pct|sm__pipe_tensor_cycles_active\.avg\.pct"), sm=col(r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed$"),
         xbar=col(r"l1tex__m_xbar2l1tex_read_bytes\.sum$"), lts=col(r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed$"), rd=col(r"dram__bytes_read\.sum$"),
         wr=col(r"dram__bytes_write\.sum$"), dram=col(r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed$"), mhz=col(r"sm__cycles_elapsed\.avg\.per_second$"),
         l1hit=col(r"l1tex__t_sector_hit_rate\.pct$"), smem=col(r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$"))


def num(r, k, scale=1.0):
    i = C[k]
    if i is None or i >= len(r) or r[i] in ("", "n/a"):
        return float("nan")
    try:
        v = float(r[i].replace(",", ""))
    except ValueError:
        return float("nan")
    u = units[i]
    if u in ("nsecond", "ns"):
        v /= 1e3
    if u in ("Mbyte",):
        v *= 1e6
    if u in ("Gbyte",):
        v *= 1e9
    if u in ("Kbyte",):
        v *= 1e3
    return v * scale


print(f"{'#':>3s} {'kernel':42s} {'grid':>6s} {'time_us':>9s} {'tensor%':>8s} {'sm%':>6s} {'L2->SM MB':>10s} {'lts%':>6s} {'dram rd MB':>11s} {'dram wr MB':>11s} {'dram%':>6s}")
n = 0
for r in body:
    if len(r) < len(hdr) or pat not in r[C["name"]]:
        continue
    name = re.sub(r"\(.*", "", r[C["name"]]).replace("void ", "").replace("vd3d::", "")
    grid = r[C["grid"]].strip("()").split(",")[0] if C["grid"] is not None else ""
    print(f"{n:3d} {name[:42]:42s} {grid:>6s} {num(r, 't'):9.1f} {num(r, 'tensor'):8.1f} {num(r, 'sm'):6.1f} {num(r, 'xbar', 1e-6):10.1f} {num(r, 'lts'):6.1f} "
          f"{num(r, 'rd', 1e-6):11.1f} {num(r, 'wr', 1e-6):11.1f} {num(r, 'dram'):6.1f}")
    n += 1
