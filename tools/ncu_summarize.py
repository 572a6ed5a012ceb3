"""Turn an `ncu --set full --import-source on` capture of rollout_kernel into the committed summaries (development
tool, runs where ncu is installed -- no GPU needed):

    python tools/ncu_summarize.py gpurun_out/r02_rollout.ncu-rep r02 65536 20

writes profiles/<tag>_rollout_kernel.md (headline metrics), profiles/<tag>_rollout_instmix.md (per-pipe utilisation,
opcode histogram, hottest source lines), profiles/<tag>_rollout_kernel_ncu_raw.csv and updates
profiles/rollout_traffic.json (what bench.py reads for roofline.traffic / executed / ncu_fp64_pipe_pct).
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True, check=True).stdout


def main():
    rep, tag, envs, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    write_traffic = "notraffic" not in sys.argv[5:]          # captures of other variants must not feed bench.py
    what = " ".join(a for a in sys.argv[5:] if a != "notraffic") or "Philox actions, reset rows carried in the device scratch"
    raw = ncu(rep, "--page", "raw", "--csv")
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    with open(os.path.join(ROOT, "profiles", f"{tag}_rollout_kernel_ncu_raw.csv"), "w") as f:
        f.write(raw)

    def val(name):
        return float(m[name][1].replace(",", ""))

    grid, block = int(val("launch__grid_size")), int(val("launch__block_size"))
    # the 512-thread CTA of the 448-env shape is 14 worker warps + 2 helper warps: the per-warp-step figures are per
    # WORKER warp (one lane = one env), with the helpers' instructions included in the totals
    helpers = 2 if (block == 512 and envs <= grid * 448) else 0
    warps = grid * (block // 32 - helpers)
    div = warps * steps
    sass = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "sass"))))
    h2 = sass[1]
    ci, cs = h2.index("Instructions Executed"), h2.index("Source")
    hist, total = collections.Counter(), 0
    for r in sass[2:]:
        if len(r) <= ci:
            continue
        src = re.sub(r"^@!?U?P\d+\s+", "", r[cs].strip())
        op = src.split()[0] if src else "?"
        n = int(r[ci])
        total += n
        hist[op] += n
    base = collections.Counter()
    for op, n in hist.items():
        base[op.split(".")[0]] += n
    fp64_ops = {k: base.get(k, 0) / div for k in ("DFMA", "DMUL", "DADD", "DSETP")}
    fp64_inst = sum(fp64_ops.values())
    flop = 2 * fp64_ops["DFMA"] + fp64_ops["DMUL"] + fp64_ops["DADD"]          # per env-step (one lane = one env)
    xu = sum(n for op, n in base.items() if op in ("MUFU", "F2F", "I2F", "F2I")) / div
    # hottest source lines
    cs_rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    per_line, cur_file, h3 = [], None, None
    for r in cs_rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            h3 = r
            ci3 = h3.index("Instructions Executed")
            continue
        if h3 and r[0] not in ("", "Function Name") and len(r) > ci3:
            try:
                per_line.append((int(r[ci3]), cur_file, int(r[0]), r[1].strip()[:90]))
            except ValueError:
                pass
    names = [
        "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    ]
    dur_us = val("gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "msecond": 1e3, "ms": 1e3,
                                              "nsecond": 1e-3}.get(m["gpu__time_duration.sum"][0], 1.0)
    fp64_pct = val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active")
    with open(os.path.join(ROOT, "profiles", f"{tag}_rollout_kernel.md"), "w") as f:
        f.write(f"# {tag} -- `rollout_kernel` (the dominant kernel of bench.py's timed region), ncu --set full, B200\n\n")
        f.write(f"Capture: `tools/ncu_rollout.sh {tag}` (`ncu --set full --clock-control none --import-source on -k "
                f"regex:rollout_kernel -s 3 -c 1 python tools/rollout_one.py {envs} {steps}`), run after the same script "
                f"had exited 0 without ncu: one launch of {envs:,} envs x {steps} steps, {what}.  Full dump: `{tag}_rollout_kernel_ncu_raw.csv`; launch list of the same "
                f"command: `{tag}_launches.csv`.\n\n| metric | value | unit |\n|---|---|---|\n")
        for nme in names:
            if nme in m:
                f.write(f"| {nme} | {m[nme][1]} | {m[nme][0]} |\n")
        f.write(f"\nPer env-step under ncu: {dur_us / steps:.2f} us for {envs:,} envs ({dur_us:.1f} us per {steps}-step "
                f"launch, cold caches, serialised).  {total / div:.0f} warp-instructions per warp and step, "
                f"{fp64_inst:.0f} of them fp64 ({fp64_ops['DFMA']:.0f} DFMA + {fp64_ops['DMUL']:.0f} DMUL + "
                f"{fp64_ops['DADD']:.0f} DADD + {fp64_ops['DSETP']:.0f} DSETP) = {flop:.0f} executed fp64 flop per "
                f"env-step; {xu:.0f} on the 16-lane XU pipe (MUFU, F2F, I2F).  fp64 pipe active {fp64_pct:.1f} %.\n")
    with open(os.path.join(ROOT, "profiles", f"{tag}_rollout_instmix.md"), "w") as f:
        f.write(f"# {tag} -- instruction mix of `rollout_kernel` ({envs:,} envs x {steps} steps, same capture as "
                f"`{tag}_rollout_kernel.md`)\n\nPer warp and env-step ({warps} warps x {steps} steps; a warp-instruction "
                f"counts once whatever its active lanes).\n\n## Pipes (ncu, % of peak while active)\n\n| pipe | % |\n|---|---|\n")
        for nme in names:
            if "pipe" in nme and nme in m:
                f.write(f"| {nme} | {m[nme][1]} |\n")
        f.write(f"| smsp__issue_active.avg.pct_of_peak_sustained_active | "
                f"{m['smsp__issue_active.avg.pct_of_peak_sustained_active'][1]} |\n")
        f.write("\n## Opcodes\n\n| opcode | per warp-step | % of all |\n|---|---|---|\n")
        for op, n in base.most_common(40):
            f.write(f"| {op} | {n / div:.1f} | {100.0 * n / total:.2f} |\n")
        f.write(f"| **all** | **{total / div:.1f}** | 100 |\n")
        f.write("\nVariants of the opcodes that are not arithmetic of the model:\n\n| opcode | per warp-step |\n|---|---|\n")
        for op, n in sorted(hist.items(), key=lambda kv: -kv[1]):
            if op.split(".")[0] in ("IMAD", "F2F", "MUFU", "LDCU", "LDL", "STL", "I2F", "BRA", "UMOV", "LOP3") and n / div >= 2:
                f.write(f"| {op} | {n / div:.1f} |\n")
        f.write("\n## Hottest source lines (all opcodes)\n\n| per warp-step | where | source |\n|---|---|---|\n")
        for n, fl, ln, src in sorted(per_line, reverse=True)[:40]:
            f.write(f"| {n / div:.1f} | {fl}:{ln} | `{src.replace('|', '/')}` |\n")
    traffic = {
        "envs": envs, "steps_per_launch": steps, "kernel": "rollout_kernel",
        "dram_bytes_read": val("dram__bytes_read.sum") * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}[m["dram__bytes_read.sum"][0]],
        "dram_bytes_write": val("dram__bytes_write.sum") * {"Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Gbyte": 1e9}[m["dram__bytes_write.sum"][0]],
        "fp64_flop_per_env_step_executed": round(flop, 1), "fp64_pipe_pct": fp64_pct,
        "xu_pipe_pct": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions_per_warp_step": round(total / div, 1), "fp64_instructions_per_warp_step": round(fp64_inst, 1),
        "source": f"profiles/{tag}_rollout_kernel_ncu_raw.csv (ncu --set full --clock-control none, one {steps}-step launch of "
                  f"tools/rollout_one.py) and the source page of the same capture (profiles/{tag}_rollout_instmix.md)",
    }
    if write_traffic:
        with open(os.path.join(ROOT, "profiles", "rollout_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main()
