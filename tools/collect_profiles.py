#!/usr/bin/env python
"""Turns the outputs of tools/measure_round2.sh (gpurun_out/<tag>_*) into the tracked files under profiles/:
  <tag>_step_<scene>.txt            per-kernel summary of the ncu --set full capture of one step (tools/ncu_summary.py, run on
                                    the GPU box; the .ncu-rep files are too big to travel)
  <tag>_lines_<scene>_<kernel>.txt  source lines with the most executed instructions / stall samples (tools/ncu_lines.py)
  <tag>_launches_human.csv          the ncu launch list of the bench command, plus <tag>_launches_human_summary.txt
  <tag>_traffic.json                DRAM bytes per launch for bench.py's roofline.traffic, per timing group of
                                    smenv_kernel_times
Usage: python tools/collect_profiles.py <tag> [short tag used for the file names under profiles/, default r02]"""
import collections
import csv
import glob
import io
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GROUP = {"joint_kernel": "joint_kernel", "joint_first_kernel": "joint_heavy_kernel",
         "joint_solve_kernel": "joint_heavy_kernel", "joint_final_kernel": "joint_heavy_kernel",
         "contact_coarse_kernel": "contact_plan_kernel", "contact_plan_kernel": "contact_plan_kernel",
         "hcontact_coarse_kernel": "contact_plan_kernel", "hcontact_plan_kernel": "contact_plan_kernel",
         "distance_plan_kernel": "distance_plan_kernel", "gjk_kernel": "gjk_kernel", "finish_kernel": "finish_kernel",
         "human_reset_kernel": "finish_kernel", "mlp_kernel": "human_policy", "human_action_kernel": "human_policy",
         "human_brake_traj_kernel": "human_brake_traj_kernel", "human_brake_plan_kernel": "human_brake_plan_kernel",
         "human_advance_kernel": "human_advance_outcome", "human_outcome_kernel": "human_advance_outcome"}


def base_name(name):
    name = name.split("(")[0].split("<")[0]
    return name.replace("void ", "").strip()


def main():
    tag = sys.argv[1]
    short = sys.argv[2] if len(sys.argv) > 2 else "r02"
    out, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    traffic = {"source": "ncu --set full --clock-control none, one step after 25 warm-up steps (tools/profile_step.py, one env "
                         "range), 65536 envs, capture {}; dram__bytes_read.sum + dram__bytes_write.sum per launch, bytes, "
                         "summed over the kernels of each timing group of smenv_kernel_times (in the Human scene the joint "
                         "kernels and the GJK kernel run twice per step: nested env and robot)".format(tag),
               "bytes_per_launch": {}, "bytes_per_kernel": {}}
    for scene in ("human", "space_bm", "space", "ball"):
        txt = os.path.join(out, "{}_step_{}.txt".format(tag, scene))
        if os.path.exists(txt) and os.path.getsize(txt):
            with open(txt) as f, open(os.path.join(prof, "{}_step_{}.txt".format(short, scene)), "w") as g:
                g.write(f.read().replace(ROOT + "/", ""))
        raw = os.path.join(out, "{}_raw_{}.csv".format(tag, scene))
        if not (os.path.exists(raw) and os.path.getsize(raw)):
            continue
        rows = list(csv.reader(io.StringIO("".join(l for l in open(raw) if not l.startswith("==")))))
        hdr, units = rows[0], rows[1]
        grp, per = collections.defaultdict(float), collections.defaultdict(float)
        for r in rows[2:]:
            d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
            b = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[k]]
                b += float(d[k].replace(",", "")) * scale
            name = base_name(d["Kernel Name"])
            per[name] += b
            grp[GROUP.get(name, name)] += b
        traffic["bytes_per_launch"][scene] = dict(grp)
        traffic["bytes_per_kernel"][scene] = dict(per)
        traffic.setdefault("bytes_per_step", {})[scene] = sum(per.values())
    if traffic["bytes_per_launch"]:
        with open(os.path.join(prof, "{}_traffic.json".format(short)), "w") as f:
            json.dump(traffic, f, indent=1)
    for p in glob.glob(os.path.join(out, "{}_lines_*.txt".format(tag))):
        if os.path.getsize(p) > 200:
            shutil.copy(p, os.path.join(prof, os.path.basename(p).replace(tag, short)))
    for name in ("step_time.txt",):
        p = os.path.join(out, "{}_{}".format(tag, name))
        if os.path.exists(p):
            shutil.copy(p, os.path.join(prof, "{}_{}".format(short, name)))
    lcsv = os.path.join(out, "{}_launches_human.csv".format(tag))
    if os.path.exists(lcsv):
        shutil.copy(lcsv, os.path.join(prof, "{}_launches_human.csv".format(short)))
        lines = [l for l in open(lcsv) if not l.startswith("==")]
        rows = list(csv.DictReader(io.StringIO("".join(lines))))
        acc = collections.OrderedDict()
        for r in rows:
            if r.get("Metric Name") != "gpu__time_duration.sum":
                continue
            v = float(r["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1e-3)
            acc.setdefault(r["Kernel Name"][:70], []).append(v)
        total = sum(sum(v) for v in acc.values())
        with open(os.path.join(prof, "{}_launches_human_summary.txt".format(short)), "w") as f:
            f.write("# ncu launch list of `python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-scenes` "
                    "(Human, 65536 envs), {}\n# per-kernel: launches, mean us, share of all profiled launches (cold-cache, "
                    "serialised: shares, not absolutes)\n".format(tag))
            for k, v in acc.items():
                f.write("{:70s} n={:4d} mean {:9.1f} us  share {:5.1f} %\n".format(k, len(v), sum(v) / len(v),
                                                                                  100 * sum(v) / total))
            step = [k for k in acc if base_name(k) in GROUP]
            st = sum(sum(acc[k]) for k in step)
            f.write("# shares within the env step only (timing groups of smenv_kernel_times):\n")
            g = collections.OrderedDict()
            for k in step:
                g[GROUP[base_name(k)]] = g.get(GROUP[base_name(k)], 0.0) + sum(acc[k])
            for k, v in g.items():
                f.write("#   {:24s} {:5.1f} %\n".format(k, 100 * v / st))


if __name__ == "__main__":
    main()
