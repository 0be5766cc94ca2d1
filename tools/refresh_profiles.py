"""Dev tool: copies the measurements of one tools/gpu_round.sh visit (gpurun_out/<tag>/) into profiles/ -- bench lines, the ncu
launch list with its per-kernel share table, text summaries of the full ncu captures, DRAM traffic / instruction counts."""
import collections
import csv
import glob
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(tag):
    src, dst = os.path.join(ROOT, "gpurun_out", tag), os.path.join(ROOT, "profiles")
    for a, b in (("bench_fnx_n1.json", "r1_bench_fnx_n1.json"), ("bench_reference_n1.json", "r1_bench_reference_n1.json"),
                 ("profile_step.txt", "r1_profile_step.txt"), ("launches.csv", "r1_launches_step_smoke_v6.csv")):
        if os.path.exists(os.path.join(src, a)):
            shutil.copy(os.path.join(src, a), os.path.join(dst, b))
    for rep in glob.glob(os.path.join(src, "ncu_*_kernel.ncu-rep")):
        k = os.path.basename(rep)[4:-len("_kernel.ncu-rep")]
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
        open(os.path.join(dst, f"r1_ncu_{k}.txt"), "w").write(out.replace(os.path.join(ROOT, "gpurun_out", tag) + "/", "").replace(f"gpurun_out/{tag}/", ""))
    rows = [r for r in csv.reader(open(os.path.join(dst, "r1_launches_step_smoke_v6.csv"))) if len(r) > 5]
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
    out = ["# r1 (final) -- one fused training iteration, workload smoke (P=200k: 20k fluid + 180k frozen background, N=28k, 5 views 512x512)",
           "# source: profiles/r1_launches_step_smoke_v6.csv = `ncu --metrics gpu__time_duration.sum --clock-control none` of tools/profile_step.py on a B200",
           "# (last iteration of the run; cold-cache, serialised launches -- in the real step the physics kernels overlap the rasterizer on a side",
           "#  stream and 4 frames run concurrently on 4 streams: compare SHARES, not absolute times)", "",
           f"{'kernel':58s} {'launches':>8s} {'us':>10s} {'share':>7s}"]
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{n[:58]:58s} {c:8d} {v:10.1f} {100 * v / tot:6.1f}%")
    out.append(f"{'TOTAL':58s} {sum(c for c, _ in agg.values()):8d} {tot:10.1f}")
    open(os.path.join(dst, "r1_step_kernel_shares.txt"), "w").write("\n".join(out) + "\n")
    t = {"smoke": {}}
    for k in ("blend_fwd", "blend_bwd"):
        s = open(os.path.join(dst, f"r1_ncu_{k}.txt")).read()
        dram = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(key + r"\s+([\d.]+) (\w+)", s)
            dram += float(m.group(1)) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m.group(2)]
        t["smoke"][k + "_kernel"] = {"dram_bytes_per_launch": int(dram),
                                     "warp_instructions_per_launch": int(float(re.search(r"smsp__inst_executed.sum\s+([\d.]+)", s).group(1))),
                                     "source": f"profiles/r1_ncu_{k}.txt (ncu --set full, one launch = 5 views, merged static+dynamic streams)"}
    json.dump(t, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
    print("\n".join(out[5:16]))
    print(out[-1])
    print(t)


if __name__ == "__main__":
    main(sys.argv[1])
