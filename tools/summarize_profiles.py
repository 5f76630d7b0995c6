"""Turn the ncu artefacts of one GPU session (gpurun_out/<tag>_*) into the tracked summaries under profiles/."""
import collections
import csv
import json
import pathlib
import subprocess
import sys

import os

ROOT = pathlib.Path(__file__).resolve().parent.parent
# SPCL_PROFILES_OUT: write the summaries somewhere else (tools/gpu_round.sh summarises ON the GPU box into gpurun_out/,
# so that only text travels back: the .ncu-rep files of one session exceed gpurun's 64 MiB return limit)
OUT = pathlib.Path(os.environ.get("SPCL_PROFILES_OUT", ROOT / "profiles"))
METRICS = [
    "gpu__time_duration.sum", "smsp__cycles_elapsed.avg.per_second", "sm__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
]


def launches(tag):
    path = ROOT / "gpurun_out" / f"{tag}_launches.csv"
    lines = [l for l in path.read_text().splitlines(True) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row["Kernel Name"], []).append(float(row["Metric Value"].replace(",", "")))
    total = sum(sum(v) for v in agg.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none  (python bench.py --steps 2 --warmup 3)",
           f"# per-launch times are cold-cache and serialised: compare SHARES.  total {total/1e3:.1f} us",
           f"{'launches':>8} {'avg_us':>10} {'share_%':>8}  kernel"]
    for k, v in agg.items():
        out.append(f"{len(v):8d} {sum(v)/len(v)/1e3:10.2f} {100*sum(v)/total:8.2f}  {k[:110]}")
    (OUT / f"{tag}_launches.txt").write_text("\n".join(out) + "\n")


def full(tag, which):
    rep = ROOT / "gpurun_out" / f"{tag}_prof_{which}.ncu-rep"
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    out = [f"# ncu --set full --clock-control none --import-source on -k regex:{"stats" if which == "fwd" else which}_kernel  ({rep.name})"]
    for r in rows[2:]:
        out.append(f"## {r[hdr.index('Kernel Name')]}")
        for m in METRICS:
            if m in hdr:
                out.append(f"{m:75s} {r[hdr.index(m)]}  {rows[1][hdr.index(m)]}")
    (OUT / f"{tag}_ncu_{which}.txt").write_text("\n".join(out) + "\n")


def dense(tag):
    """first launch of each dense front-end kernel in the capture (32 x 128 x 224 x 224 -> 32 x 32)."""
    reps = [r for r in (ROOT / "gpurun_out" / f"{tag}_prof_dense{sfx}.ncu-rep" for sfx in ("", "_fwd", "_bwd")) if r.exists()]
    if not reps:
        return
    out = ["# ncu --set full --clock-control none --import-source on -k regex:pool_rows_{fwd,bwd} -s 3 -c 1  python tools/gpu_dense_bench.py",
           "# shape 32 x 128 x 224 x 224 -> 32 x 32; algorithmic bytes: fwd 4*B*C*H*W + 4*B*ph*pw*C = 839 MB; bwd 4*B*C*H*W written + 17 MB of rows read"]
    seen = set()
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr = rows[0]
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            out.append(f"## {name}")
            for m in METRICS:
                if m in hdr:
                    out.append(f"{m:75s} {r[hdr.index(m)]}  {rows[1][hdr.index(m)]}")
    (OUT / f"{tag}_ncu_dense.txt").write_text("\n".join(out) + "\n")


def aux(tag):
    """every kernel of the tools/gpu_aux_bench.py capture (first launch per kernel name and grid)."""
    rep = ROOT / "gpurun_out" / f"{tag}_prof_aux.ncu-rep"
    if not rep.exists():
        return
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    out = ["# ncu --set full --clock-control none -k regex:l2norm|prepare|raw_bwd|transpose  python tools/gpu_aux_bench.py",
           "# first launch per (kernel, grid): cfg3 shapes ([32768, 128] rows / [32, 128, 32, 32]) then cfg4 shapes"]
    seen = set()
    for r in rows[2:]:
        key = (r[hdr.index("Kernel Name")], r[hdr.index("launch__grid_size")] if "launch__grid_size" in hdr else "")
        if key in seen:
            continue
        seen.add(key)
        out.append(f"## {key[0]}  grid {key[1]}")
        for m in METRICS:
            if m in hdr:
                out.append(f"{m:75s} {r[hdr.index(m)]}  {rows[1][hdr.index(m)]}")
    (OUT / f"{tag}_ncu_aux.txt").write_text("\n".join(out) + "\n")


if __name__ == "__main__":
    tag = sys.argv[1]
    OUT.mkdir(parents=True, exist_ok=True)
    launches(tag)
    for which in ("fwd", "bwd"):
        full(tag, which)
    dense(tag)
    aux(tag)
    for name in ("bench.json", "bench_ref.json", "smoke.log", "pytest_gpu.log"):
        src = ROOT / "gpurun_out" / f"{tag}_{name}"
        if src.exists():
            (OUT / f"{tag}_{name}").write_text(src.read_text())
    print("wrote", sorted(p.name for p in OUT.glob(f"{tag}_*")))
