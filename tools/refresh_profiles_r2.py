"""Dev tool (no GPU needed): copies the round-2 measurements from gpurun_out/ into profiles/ -- ncu launch lists with per-kernel share
tables, text summaries of the `ncu --set full` captures of the blend kernels for every workload, and profiles/traffic.json (DRAM
bytes + warp instructions per launch, read by bench.py for `roofline.traffic` / `roofline.issue_rate`)."""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "profiles")
DESC = {"smoke": "smoke (P=200k: 20k fluid + 180k frozen background, C=3, N=28k, 5 views 512x512)",
        "scalar": "scalar (P=V=150k fluid, C=1, N=28k, 5 views 512x512)", "c2": "c2 (P=50k fluid, C=3, N=28k, 5 views 400x400)",
        "ball": "ball (P=300k: 30k fluid + 270k frozen background incl. the 30k-Gaussian ball, C=3, N=28k, 5 views 512x512)"}


def shares(wl, src):
    rows = [r for r in csv.reader(open(os.path.join(src, "launches.csv"))) if len(r) > 5]
    for i, r in enumerate(rows):
        if r[0] == "ID":
            hdr, rows = r, rows[i + 1:]
            break
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    names = [(re.sub(r"[<(].*", "", r[ik]).replace("void ", ""), float(r[iv].replace(",", ""))) for r in rows if r[im] == "gpu__time_duration.sum"]
    last = names[[i for i, (n, _) in enumerate(names) if "next_tick" in n][-1]:]
    agg = collections.OrderedDict()
    for n, v in last:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v / 1000
    tot = sum(v for _, v in agg.values())
    out = [f"# r2 -- one fused training iteration, workload {DESC[wl]}",
           f"# source: profiles/r2_launches_step_{wl}.csv = `ncu --metrics gpu__time_duration.sum --clock-control none` of tools/profile_step.py on a B200",
           "# (last iteration of the run; cold-cache, serialised launches -- in the real step the physics kernels overlap the rasterizer on a side",
           "#  stream and 4 frames run concurrently on 4 streams: compare SHARES, not absolute times)", "",
           f"{'kernel':58s} {'launches':>8s} {'us':>10s} {'share':>7s}"]
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{n[:58]:58s} {c:8d} {v:10.1f} {100 * v / tot:6.1f}%")
    out.append(f"{'TOTAL':58s} {sum(c for c, _ in agg.values()):8d} {tot:10.1f}")
    open(os.path.join(DST, f"r2_step_kernel_shares_{wl}.txt"), "w").write("\n".join(out) + "\n")
    return out


def main(tag="r2e"):
    t = {}
    for wl in ("smoke", "scalar", "c2", "ball"):
        src = os.path.join(ROOT, "gpurun_out", tag, wl)
        if not os.path.isdir(src):
            continue
        shutil.copy(os.path.join(src, "launches.csv"), os.path.join(DST, f"r2_launches_step_{wl}.csv"))
        if os.path.exists(os.path.join(src, "profile_step.txt")):
            shutil.copy(os.path.join(src, "profile_step.txt"), os.path.join(DST, f"r2_profile_step_{wl}.txt"))
        out = shares(wl, src)
        print("\n".join(out[5:12]), "\n", out[-1])
        t[wl] = {}
        for rep in sorted(os.listdir(src)):
            if not rep.endswith(".ncu-rep"):
                continue
            k = rep[4:-len("_kernel.ncu-rep")]
            s = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), os.path.join(src, rep)], capture_output=True, text=True).stdout
            s = s.replace(src + "/", "")
            open(os.path.join(DST, f"r2_ncu_{wl}_{k}.txt"), "w").write(f"# workload {DESC[wl]}; ncu --set full --clock-control none, one warm launch of tools/profile_step.py\n" + s)
            if k in ("blend_fwd", "blend_bwd"):
                dram = 0.0
                for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    m = re.search(key + r"\s+([\d.]+) (\w+)", s)
                    dram += float(m.group(1)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
                t[wl][k + "_kernel"] = {"dram_bytes_per_launch": int(dram),
                                        "warp_instructions_per_launch": int(float(re.search(r"smsp__inst_executed.sum\s+([\d.]+)", s).group(1))),
                                        "capture": f"profiles/r2_ncu_{wl}_{k}.txt (ncu --set full of this workload at 1 GPU, one launch = 5 views)"}
    json.dump(t, open(os.path.join(DST, "traffic.json"), "w"), indent=1)
    print(json.dumps(t)[:400])


if __name__ == "__main__":
    main(*sys.argv[1:])
