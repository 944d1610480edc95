"""Developer aid: turns the scratch ncu outputs under gpurun_out/ into the committed summaries
under profiles/ (launch list of the bench command, per-kernel table of the --set full captures).
usage: python tests/prof_summarize.py"""
import collections
import csv
import json
import re
import shutil
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT = ROOT / "gpurun_out"
PROF = ROOT / "profiles"


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"^bce::", "", n)
    return re.sub(r"\(.*$", "", n)


def launch_list():
    shutil.copy(OUT / "launches_r2_1gb.csv", PROF / "r2_launches_enwik1GB.csv")
    lines = [l for l in open(PROF / "r2_launches_enwik1GB.csv") if not l.startswith("==")]
    r = list(csv.reader(lines))
    hdr = r[0]
    ci = {h: i for i, h in enumerate(hdr)}
    data = [x for x in r[1:] if len(x) == len(hdr)]
    names = [x[ci["Kernel Name"]] for x in data]
    vals = [float(x[ci["Metric Value"]].replace(",", "")) for x in data]
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0}[data[0][ci["Metric Unit"]]]
    starts = [i for i, n in enumerate(names) if n.startswith("pad_cyclic")] + [len(names)]

    def summarize(a, b):
        agg = collections.OrderedDict()
        for n, v in zip(names[a:b], vals[a:b]):
            e = agg.setdefault(short(n), [0, 0.0])
            e[0] += 1
            e[1] += v * scale
        return agg

    bench = json.loads(open(OUT / "bench_plain_for_ncu.json").read().strip().splitlines()[-1])
    out = ["# Round 2: ncu launch list of the bench command, enwik-shaped 1 GB\n",
           "Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cli`",
           f"(raw list: `r2_launches_enwik1GB.csv`, {len(data)} launches = 4 resident passes + 3 passes through the host-buffer API).",
           "Times under ncu are serialised and cold-cache; what must agree with the bench is each kernel's SHARE of the step.\n",
           f"Same command without ncu (same box, just before): value {bench['value']:.0f} MB/s, {bench['ms_per_step']:.1f} ms/step, "
           f"e2e {bench['e2e']['value']:.0f} MB/s; stage_ms {json.dumps({k: round(v, 1) for k, v in bench['stage_ms'].items()})}\n"]
    for title, (a, b) in (("Timed resident step (4th pass)", (starts[3], starts[4])),
                          ("Timed host-buffer step (7th pass)", (starts[6], starts[7]))):
        agg = summarize(a, b)
        tot = sum(v[1] for v in agg.values())
        out.append(f"## {title}: {b - a} launches, {tot:.1f} ms of kernel time under ncu\n")
        out += ["| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
        for k, (c, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            out.append(f"| `{k}` | {c} | {ms:.2f} | {100 * ms / tot:.1f} % |")
        out.append("")
    st, tot = bench["stage_ms"], bench["ms_per_step"]
    out += ["## Shares in the un-profiled bench step (CUDA events)\n", "| stage | ms | share |", "|---|---:|---:|"]
    for k in ("ms_radix", "ms_cse", "ms_rerank", "ms_rekey", "ms_bwt_gather", "ms_wavelet", "ms_pack"):
        out.append(f"| {k[3:]} | {st[k]:.1f} | {100 * st[k] / tot:.1f} % |")
    out.append("\n(`radix` = the `radix_onesweep_kernel` launches of the sorts; `rerank` = `rerank_kernel` + `scatter_ranks_kernel`; `cse` = all `cse_*` kernels.)")
    (PROF / "r2_launches_enwik1GB.md").write_text("\n".join(out) + "\n")


def kernel_tables():
    stall_prefix = "smsp__average_warps_issue_stalled_"
    out = ["# Round 2: `ncu --set full --clock-control none` captures, enwik-shaped 100 MB, packed (CODER) emission",
           "", "Command: `ncu --set full --clock-control none --import-source on -k regex:<kernels> -c N python tests/gpu_trace.py enwik-shaped 100000000 2`",
           "(first pass of a fresh process, `BCE_TRACE_WARM=0`; the reports `gpurun_out/prof_r2_bwt.ncu-rep`, `prof_r2_cse.ncu-rep` are scratch, not committed).",
           "One row per captured launch, values from `ncu -i <rep> --page raw --csv`. DRAM GB = dram__bytes_read.sum + dram__bytes_write.sum; "
           "ld/st s/r = L1 sectors per global load/store request (32 = every lane its own sector).", ""]
    for f, title in (("prof_r2_bwt", "Stage A kernels (first 26 launches: round 0 sort, re-rank, scatter, tile sort of round 1 ...)"),
                     ("prof_r2_cse", "Stage A tail + stage B kernels (binned BWT, wavelet passes, level loop)")):
        csv_path = OUT / f"{f}_raw.csv"                      # written on the GPU box: the .ncu-rep files are too large to bring back
        raw = csv_path.read_text() if csv_path.exists() else subprocess.run(
            ["ncu", "-i", str(OUT / f"{f}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        h, units = rows[0], rows[1]
        ci = {k: i for i, k in enumerate(h)}
        out.append(f"## {title}\n")
        out.append("| # | kernel | grid x block | regs | time | DRAM GB | DRAM % | L1 % | L2 % | ld s/r | st s/r | warps active % | issue active % | warp instr (M) | top stalls (warps per issue cycle) |")
        out.append("|---:|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|")
        for n, v in enumerate(rows[2:]):
            def g(k):
                try:
                    return float(v[ci[k]].replace(",", ""))
                except Exception:
                    return float("nan")

            def tobytes(k):
                return g(k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(units[ci[k]], 1)
            dr = (tobytes("dram__bytes_read.sum") + tobytes("dram__bytes_write.sum")) / 1e9
            stalls = []
            for k, i in ci.items():
                if k.startswith(stall_prefix) and "not_issued" not in k and k.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(v[i]), k[len(stall_prefix):-len("_per_issue_active.ratio")]))
                    except Exception:
                        pass
            stalls.sort(reverse=True)
            st = ", ".join(f"{nm} {x:.1f}" for x, nm in stalls[:3])
            ldr = g("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") / max(g("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), 1)
            strq = g("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum") / max(g("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum"), 1)
            out.append(f"| {n} | `{short(v[ci['Kernel Name']])}` | {int(g('launch__grid_size'))} x {int(g('launch__block_size'))} | "
                       f"{int(g('launch__registers_per_thread'))} | {g('gpu__time_duration.sum'):.3f} {units[ci['gpu__time_duration.sum']]} | {dr:.3f} | "
                       f"{g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {g('l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                       f"{g('lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {ldr:.1f} | {strq:.1f} | "
                       f"{g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
                       f"{g('smsp__inst_executed.sum') / 1e6:.1f} | {st} |")
        out.append("")
    (PROF / "r2_ncu_kernels.md").write_text("\n".join(out) + "\n")


if __name__ == "__main__":
    launch_list()
    kernel_tables()
    print("profiles/ updated")
