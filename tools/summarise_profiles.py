"""Summarise gpurun_out/launches.csv and *.ncu-rep into small text files under profiles/ (run in the build container)."""
import csv, io, json, os, subprocess, sys, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
prefix = sys.argv[2] if len(sys.argv) > 2 else ""          # gpurun_out/<prefix>launches.csv, <prefix>prof_*.ncu-rep
out_dir = os.path.join(ROOT, "profiles"); os.makedirs(out_dir, exist_ok=True)

def short(name):
    name = re.sub(r"\(.*", "", name).replace("void ", "")
    return name.split("::")[-1][:60] if "at::" not in name else "torch:" + name.split("<")[0].split("::")[-1]

# ---- launch list: per-kernel share of one planning step (the timed region replays a CUDA graph of exactly these kernels)
rows = []
with open(os.path.join(ROOT, "gpurun_out", prefix + "launches.csv")) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(io.StringIO("".join(lines))):
    try:
        rows.append((int(r["ID"]), r["Kernel Name"], float(r["Metric Value"])))
    except Exception:
        pass
# one eager planning step = from a step_tick_kernel to the next adam_clamp_kernel
ticks = [i for i, r in enumerate(rows) if "step_tick_kernel" in r[1]]
adams = [i for i, r in enumerate(rows) if "adam_clamp_kernel" in r[1]]
seg = None
for t in ticks:
    nxt = [a for a in adams if a > t]
    if nxt:
        seg = rows[t:nxt[0] + 1]
        break
with open(os.path.join(out_dir, f"{tag}_launches_one_step.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none ... python bench.py --steps 2 --warmup 3\n")
    f.write("# kernels of ONE planning step (B=64, T=200, math=bf16), cold-cache serialised durations: compare SHARES\n")
    if seg:
        tot = sum(r[2] for r in seg)
        agg = collections.OrderedDict()
        for _, n, v in seg:
            k = short(n); a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
        f.write(f"# launches in the step: {len(seg)}   total {tot/1e6:.3f} ms\n")
        f.write(f"{'kernel':62s} {'n':>4s} {'ms':>9s} {'share':>7s}\n")
        for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:62s} {n:4d} {v/1e6:9.3f} {100*v/tot:6.1f}%\n")
        f.write("\n# in launch order\n")
        for i, n, v in seg:
            f.write(f"{i:5d} {short(n):62s} {v/1e3:10.1f} us\n")
print(open(os.path.join(out_dir, f"{tag}_launches_one_step.txt")).read()[:3000])

# ---- full captures
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_uniform", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__cluster", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__shared_mem_per_block_dynamic",
           "smsp__inst_executed.sum", "sm__inst_executed_pipe_tc", "tensor", "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__throughput.avg",
           "sm__cycles_active.avg"]
for rep in ("prof_fwd", "prof_bwd", "prof_gemm", "prof_gemm2", "prof_gategemm", "prof_elem"):
    path = os.path.join(ROOT, "gpurun_out", prefix + rep + ".ncu-rep")
    if not os.path.exists(path):
        continue
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(txt)))
    if len(rd) < 3:
        continue
    hdr, units = rd[0], rd[1]
    with open(os.path.join(out_dir, f"{tag}_ncu_{rep[5:]}.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on ({rep}.ncu-rep), bench.py B=64 T=200 math=bf16\n")
        for row in rd[2:]:
            d = dict(zip(hdr, row))
            f.write(f"\n## {short(d.get('Kernel Name',''))}  grid {d.get('Grid Size','')} block {d.get('Block Size','')}\n")
            for h, u, v in zip(hdr, units, row):
                if any(m in h for m in METRICS):
                    f.write(f"{h:80s} {v:>20s} {u}\n")
    print("wrote", rep)

# ---- DRAM traffic of the recurrent kernels per (word, cell step), for bench.py's roofline.traffic
def _metric(path, name):
    for line in open(path):
        if line.startswith(name + " "):
            parts = line.split()
            v, unit = float(parts[1]), parts[2]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9}.get(unit, 1)
    return None
traffic = {}
for kind in ("fwd", "bwd"):
    path = os.path.join(out_dir, f"{tag}_ncu_{kind}.txt")
    if os.path.exists(path):
        rd, wr, dur = _metric(path, "dram__bytes_read.sum"), _metric(path, "dram__bytes_write.sum"), _metric(path, "gpu__time_duration.sum")
        if rd is not None and wr is not None and dur:
            # the captured launch is one layer of the bench workload (B = 64, T = 200): the forward model (200 steps; forward: the
            # 2-quarter fused layout <2, 1, 1, 2> or, without the dense pipeline, <1, 1, 1, 1>; backward: > 500 us) or an embedder
            # layer (100 steps)
            head = open(path).read(400)
            if kind == "fwd":
                steps = 200 if ("fwd2_kernel<2, 1, 1, 2>" in head or "fwd2_kernel<1, 1, 1, 1>" in head) else 100
            else:
                steps = 200 if dur > 500e-6 else 100
            traffic[kind] = {"dram_bytes_per_word_step": (rd + wr) / (64 * steps), "captured_steps": steps, "captured_us": dur * 1e6,
                             "source": f"profiles/{tag}_ncu_{kind}.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
json.dump(traffic, open(os.path.join(out_dir, "ncu_traffic.json"), "w"), indent=1)
print("traffic", traffic)
