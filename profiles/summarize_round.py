"""Builds profiles/<round>_summary.md from the artifacts a GPU run leaves under gpurun_out/:

  <round>_bench.json            one `python bench.py` line (not under a profiler)
  <round>_bench_reference.json  one `python bench.py --impl reference` line
  <round>_launches.csv          ncu --metrics gpu__time_duration.sum --clock-control none (launch list)
  <round>_k1_full.ncu-rep       ncu --set full --import-source on of containerAndCountKernel
  <round>_other_full.ncu-rep    the same for the small kernels

and copies the small text artifacts next to it (the .ncu-rep files stay in gpurun_out/, which is scratch).
Usage: python profiles/summarize_round.py r1
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROFILES = os.path.join(ROOT, "profiles")


def export(report, page):
    path = os.path.join(OUT, report)
    return list(csv.reader(subprocess.run(
        ["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True, check=True).stdout.splitlines()))


def launch_table(tag):
    rows = list(csv.reader(open(os.path.join(OUT, f"{tag}_launches.csv"))))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    header = rows[start]
    name_at, value_at = header.index("Kernel Name"), header.index("Metric Value")
    per_kernel = collections.OrderedDict()
    for row in rows[start + 1:]:
        if len(row) > value_at:
            name = row[name_at].split("(")[0].replace("void ", "").replace("silo::<unnamed>::", "")
            per_kernel.setdefault(name, []).append(float(row[value_at].replace(",", "")) / 1000.0)
    step_kernels = [k for k in per_kernel if len(per_kernel[k]) > 2]
    step_total = sum(sum(per_kernel[k]) / len(per_kernel[k]) for k in step_kernels)
    lines = ["| kernel | launches | mean µs (cold, serialised) | share of a step |", "|---|---|---|---|"]
    for name, values in per_kernel.items():
        mean = sum(values) / len(values)
        share = f"{100 * mean / step_total:.1f} %" if name in step_kernels else "setup, once"
        lines.append(f"| `{name[:70]}` | {len(values)} | {mean:.2f} | {share} |")
    return "\n".join(lines), step_total, per_kernel


RAW_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def raw_tables(report):
    rows = export(report, "raw")
    header, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        lines = [f"`{row[header.index('Kernel Name')][:90]}`", "", "| metric | value | unit |", "|---|---|---|"]
        values = {}
        for metric in RAW_METRICS:
            if metric in header:
                values[metric] = row[header.index(metric)]
                lines.append(f"| {metric} | {row[header.index(metric)]} | {units[header.index(metric)]} |")
        out.append(("\n".join(lines), values, units))
    return out


def source_regions(report):
    rows = export(report, "source")
    body = rows[1:]
    header, data = body[0], body[1:]
    source_at, executed_at, samples_at = header.index("Source"), header.index("Instructions Executed"), header.index("# Samples")
    groups = []
    for i, row in enumerate(data):
        if len(row) <= executed_at or not row[executed_at]:
            continue
        executed, samples = int(row[executed_at]), int(row[samples_at])
        if groups and groups[-1]["executed"] == executed and groups[-1]["end"] == i - 1:
            group = groups[-1]
            group["end"] = i
            group["n"] += 1
            group["samples"] += samples
        else:
            groups.append({"start": i, "end": i, "executed": executed, "n": 1, "samples": samples, "first": row[source_at].strip()})
    total = sum(g["executed"] * g["n"] for g in groups)
    total_samples = sum(g["samples"] for g in groups)
    lines = [f"{total} warp instructions, {total_samples} stall samples. Regions of consecutive SASS instructions with the same "
             "execution count (>= 0.8 % of the instructions or >= 1.5 % of the samples):", "",
             "| SASS lines | instrs | executions | warp instrs (M) | share | samples | first instruction |", "|---|---|---|---|---|---|---|"]
    for g in groups:
        share = 100.0 * g["executed"] * g["n"] / total
        if share >= 0.8 or g["samples"] >= 0.015 * total_samples:
            lines.append(f"| {g['start']}-{g['end']} | {g['n']} | {g['executed']} | {g['executed'] * g['n'] / 1e6:.2f} | {share:.1f} % | "
                         f"{g['samples']} | `{g['first'][:60]}` |")
    stall_columns = [i for i, h in enumerate(header) if h.startswith("stall_")]
    stalls = {}
    for row in data:
        for i in stall_columns:
            if i < len(row) and row[i]:
                stalls[header[i]] = stalls.get(header[i], 0) + int(row[i])
    stall_total = sum(stalls.values())
    stall_line = ", ".join(f"{k[6:]} {100.0 * v / stall_total:.1f} %" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:10])
    return "\n".join(lines), stall_line, total


def extras(tag):
    """Round 2 on: the other workloads' bench lines, the threshold sweep kernel, the multi-GPU runs and the sanitizer logs
    (each only if round_capture.sh / scripts/r2_scale.sh left the artifact)."""
    text = ""
    lines = []
    for workload in ("nof", "aa", "reads", "cooc"):
        path = os.path.join(OUT, f"{tag}_bench_{workload}.json")
        if os.path.exists(path):
            shutil.copy(path, os.path.join(PROFILES, f"{tag}_bench_{workload}.json"))
            line = json.load(open(path))
            parity = line.get("parity", {})
            lines.append(f"| `--workload {workload}` | {line['config']['workload'][:110]} | {line['value']:.4g} {line['unit']} | "
                         f"{line['e2e']['value']:.4g} ({line['e2e']['ms_per_step']:.3f} ms/step) | {line['cpu_baseline']['value']:.4g} "
                         f"({line['cpu_baseline']['cores']} core) | {', '.join(f'{k}: {v}' for k, v in parity.items() if isinstance(v, bool))} |")
    if lines:
        text += ("\n## the other workloads of BASELINE.json (`profiles/" + tag + "_bench_<workload>.json`, not under a profiler)\n\n"
                 "| command | workload | value | e2e | cpu_baseline | parity at full size |\n|---|---|---|---|---|---|\n" + "\n".join(lines) + "\n")
    if os.path.exists(os.path.join(OUT, f"{tag}_sweep_full.ncu-rep")):
        sweep = raw_tables(f"{tag}_sweep_full.ncu-rep")
        text += ("\n## thresholdSweepKernel (configs[2]: one nucleotideMutationProfile filter, distance 5, 10 M rows), "
                 "`ncu --set full` of `profiles/config3_ncu_target.py`\n\n" + sweep[0][0] + "\n")
    path = os.path.join(OUT, f"{tag}_config3_launches.csv")
    if os.path.exists(path):
        shutil.copy(path, os.path.join(PROFILES, f"{tag}_config3_launches.csv"))
        rows = list(csv.reader(open(path)))
        start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        header = rows[start]
        text += "\nLaunch list of the two filters of that script (`profiles/" + tag + "_config3_launches.csv`, cold, serialised):\n\n| kernel | µs |\n|---|---|\n"
        for row in rows[start + 1:]:
            if len(row) > header.index("Metric Value"):
                text += f"| `{row[header.index('Kernel Name')].split('(')[0][:80]}` | {float(row[header.index('Metric Value')].replace(',', '')) / 1000.0:.2f} |\n"
    path = os.path.join(OUT, f"{tag}_cooc_launches.csv")
    if os.path.exists(path):
        shutil.copy(path, os.path.join(PROFILES, f"{tag}_cooc_launches.csv"))
        rows = list(csv.reader(open(path)))
        start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
        header = rows[start]
        name_at, metric_at, value_at = header.index("Kernel Name"), header.index("Metric Name"), header.index("Metric Value")
        per_kernel = collections.OrderedDict()
        for row in rows[start + 1:]:
            if len(row) > value_at:
                kernel = row[name_at].split("(")[0].replace("unnamed>::", "")
                per_kernel.setdefault(kernel, collections.defaultdict(list))[row[metric_at]].append(float(row[value_at].replace(",", "")))
        text += ("\n## the co-occurrence kernels (`bench.py --workload cooc`, 10 M rows x 6 positions; `profiles/" + tag + "_cooc_launches.csv`: "
                 "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`, cold, serialised)\n\n"
                 "| kernel | launches | mean µs | DRAM read MB | DRAM write MB | GB/s | of the measured peak |\n|---|---|---|---|---|---|---|\n")
        peak = json.load(open(os.path.join(OUT, f"{tag}_bench.json")))["roofline"]["peak"]
        for kernel, metrics in per_kernel.items():
            n = len(metrics["gpu__time_duration.sum"])
            micros = sum(metrics["gpu__time_duration.sum"]) / n / 1000.0
            read_mb = sum(metrics["dram__bytes_read.sum"]) / n / 1e6
            write_mb = sum(metrics["dram__bytes_write.sum"]) / n / 1e6
            gbs = (read_mb + write_mb) / micros * 1e3
            text += f"| `{kernel}` | {n} | {micros:.1f} | {read_mb:.1f} | {write_mb:.1f} | {gbs:.0f} | {gbs / peak:.3f} |\n"
    scale = []
    for n in (1, 2, 4, 8):
        for reduce in ("", "_peer", "_nccl"):
            path = os.path.join(OUT, f"{tag}_scale_n{n}{reduce}.json")
            if os.path.exists(path):
                shutil.copy(path, os.path.join(PROFILES, os.path.basename(path)))
                line = json.loads([l for l in open(path) if l.startswith("{")][-1])
                scale.append((n, reduce.strip("_") or "-", line))
    if scale:
        base = next(line for n, _, line in scale if n == 1)
        text += ("\n## one process per GPU, 1 -> 8 B200 (`profiles/scripts/" + tag + "_scale.sh`, weak scaling: 10 M rows per GPU)\n\n"
                 "| GPUs | reduce | µs/step (device, max over ranks) | value | efficiency | e2e µs/step | e2e value | e2e efficiency |\n|---|---|---|---|---|---|---|---|\n")
        for n, reduce, line in scale:
            text += (f"| {n} | {reduce} | {line['ms_per_step'] * 1000:.1f} | {line['value']:.4g} | {line['value'] / (n * base['value']):.3f} | "
                     f"{line['e2e']['ms_per_step'] * 1000:.1f} | {line['e2e']['value']:.4g} | {line['e2e']['value'] / (n * base['e2e']['value']):.3f} |\n")
        text += ("\n`peer`: the library's shard group (finalize kernels store their rows into the root's memory over NVLink, the root sums "
                 "them in a kernel); `nccl`: torch.distributed all_reduce of the count arrays between the container kernel and the finalize kernel.\n")
    sanitizer = []
    for name in sorted(os.listdir(OUT)):
        if name.startswith(f"{tag}_sanitizer_") and name.endswith(".log") and "pytest" not in name:
            shutil.copy(os.path.join(OUT, name), os.path.join(PROFILES, name))
            summary = [l.strip() for l in open(os.path.join(OUT, name)) if "SUMMARY" in l]
            pytest_log = os.path.join(OUT, name.replace(".log", "_pytest.log"))
            tests = [l.strip() for l in open(pytest_log) if " passed" in l or " failed" in l] if os.path.exists(pytest_log) else []
            sanitizer.append(f"| `{name}` | {summary[-1] if summary else 'no summary line'} | {tests[-1] if tests else ''} |")
    if sanitizer:
        text += "\n## compute-sanitizer over the GPU suites (`profiles/round_capture.sh`)\n\n| log | result | pytest |\n|---|---|---|\n" + "\n".join(sanitizer) + "\n"
    return text


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    bench = json.load(open(os.path.join(OUT, f"{tag}_bench.json")))
    reference = json.load(open(os.path.join(OUT, f"{tag}_bench_reference.json")))
    for name in (f"{tag}_bench.json", f"{tag}_bench_reference.json", f"{tag}_launches.csv"):
        shutil.copy(os.path.join(OUT, name), os.path.join(PROFILES, name))
    table, step_total, per_kernel = launch_table(tag)
    k1_tables = raw_tables(f"{tag}_k1_full.ncu-rep")
    k1_table, k1_values, _ = k1_tables[0]
    regions, stall_line, _ = source_regions(f"{tag}_k1_full.ncu-rep")
    others = raw_tables(f"{tag}_other_full.ncu-rep")
    roofline = bench["roofline"]
    dram = (float(k1_values["dram__bytes_read.sum"]) + float(k1_values["dram__bytes_write.sum"])) * 1e6
    with open(os.path.join(PROFILES, f"{tag}_k1_traffic.json"), "w") as out:
        json.dump({"kernel": "containerAndCountKernel", "dram_bytes_per_launch": dram,
                   "source": f"ncu --set full capture {tag}_k1_full.ncu-rep: dram__bytes_read.sum + dram__bytes_write.sum",
                   "algorithmic_bytes_per_launch": roofline["algorithmic_bytes_per_launch"]}, out, indent=1)
    k1_name = next(k for k in per_kernel if k.startswith("containerAndCountKernel"))
    k1_share_ncu = sum(per_kernel[k1_name]) / len(per_kernel[k1_name]) / step_total
    coverage = [k for k in per_kernel if "coverageDiffKernel" in k]
    coverage_mean = sum(per_kernel[coverage[0]]) / len(per_kernel[coverage[0]]) if coverage else 0.0
    k1_share_ncu_without_coverage = sum(per_kernel[k1_name]) / len(per_kernel[k1_name]) / (step_total - coverage_mean)
    k1_share_live = roofline["kernel_ms"] / bench["ms_per_step"]
    text = f"""# {tag}: ncu summary (B200, sm_100a)

All numbers on this page come from `gpurun` runs of this repository on one B200; the `.ncu-rep` files stay
in `gpurun_out/` (scratch), this page and the small text artifacts next to it are what is committed.
Regenerate with `python profiles/summarize_round.py {tag}`.

## bench line (not under a profiler) -- `profiles/{tag}_bench.json`

| | |
|---|---|
| workload | {bench['config']['workload']} |
| value (inputs resident in HBM) | {bench['value']:.4g} {bench['unit']}, {bench['ms_per_step'] * 1000:.1f} µs/step |
| e2e (host API, H2D + D2H inside) | {bench['e2e']['value']:.4g} {bench['unit']}, {bench['e2e']['ms_per_step'] * 1000:.1f} µs/step, h2d {bench['e2e']['h2d_bytes_per_step']} B, d2h {bench['e2e']['d2h_bytes_per_step']} B |
| dominant kernel | `{roofline['kernel']}`: {roofline['kernel_ms'] * 1000:.1f} µs/launch over {roofline['timed_launches']} launches (CUDA events on the launching stream) |
| algorithmic bytes / launch | {roofline['algorithmic_bytes_per_launch']} |
| achieved | {roofline['achieved']:.0f} GB/s = {100 * roofline['frac']:.1f} % of {roofline['peak']} GB/s ({roofline['peak_source']}) |
| DRAM traffic / launch (ncu) | {dram:.0f} B = {dram / roofline['algorithmic_bytes_per_launch']:.3f} x algorithmic |
| clocks | {bench['clocks']} |
| cpu_baseline (oracle port, 1 core) | {bench['cpu_baseline']['value']:.4g} {bench['unit']} -- {bench['cpu_baseline']['sample']} |
| reference arm (oracle port, all host cores) | {reference['value']:.4g} {reference['unit']} -- {reference['cpu_baseline']['sample']} |

## launch list -- `profiles/{tag}_launches.csv`

`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 3 --warmup 3 --skip-cpu-baseline`
(per-launch times are cold-cache and serialised; only the SHARES are comparable with the live run):

{table}

The dominant kernel's share of a step: {100 * k1_share_ncu:.1f} % in the launch list, {100 * k1_share_live:.1f} % in the live run
({roofline['kernel_ms'] * 1000:.1f} of {bench['ms_per_step'] * 1000:.1f} µs; in the live step the coverage kernel runs BESIDE the
container kernel on a second stream, so the serialised launch list counts it on top: without it the shares are
{100 * k1_share_ncu_without_coverage:.1f} % and {100 * k1_share_live:.1f} %).

## containerAndCountKernel, `ncu --set full --clock-control none --import-source on`

{k1_table}

No tensor-core activity (bit intersection, HBM-bound by design). `UBLKCP` (1-D TMA bulk copies) in the SASS.

### where the instructions and the stalls are

{regions}

Stall reasons over all samples: {stall_line}.

## the small kernels (same capture settings)

""" + "\n\n".join(t for t, _, _ in others) + "\n"
    text += extras(tag)
    with open(os.path.join(PROFILES, f"{tag}_summary.md"), "w") as out:
        out.write(text)
    print(f"wrote profiles/{tag}_summary.md, profiles/{tag}_k1_traffic.json")


if __name__ == "__main__":
    main()
