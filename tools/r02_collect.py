"""Turns the files a GPU call left under gpurun_out/ into the evidence kept under profiles/ (round 2)."""
import csv
import glob
import json
import os
import shutil
import statistics as st
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json_line(path):
    for ln in reversed(open(path).read().strip().splitlines()):
        if ln.startswith("{"):
            return json.loads(ln)
    raise ValueError(path)


def copy_bench(src, dst):
    if os.path.exists(os.path.join(G, src)):
        d = last_json_line(os.path.join(G, src))
        json.dump(d, open(os.path.join(P, dst), "w"))
        print(dst, "%.2f ms/step" % d.get("ms_per_step", float("nan")))


def traffic_from_csv(src, dst):
    path = os.path.join(G, src)
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    idi = hdr.index("ID")
    per = {}
    for r in rows[1:]:
        per.setdefault(r[idi], {"kernel": r[ki]})[r[mi]] = float(r[vi].replace(",", ""))
    kernels = {}
    for v in per.values():
        name = v["kernel"].replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("void ", "").replace("thz::", "")
        name = name.split("(")[0].split("<")[0].strip()
        rd, wr = v.get("dram__bytes_read.sum", 0.0), v.get("dram__bytes_write.sum", 0.0)
        e = kernels.setdefault(name, {"launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "gpu_time_ms_under_ncu": 0.0})
        e["launches"] += 1
        e["dram_bytes_read"] += rd
        e["dram_bytes_write"] += wr
        e["gpu_time_ms_under_ncu"] += v.get("gpu__time_duration.sum", 0.0) / 1e6
    for e in kernels.values():
        n = e["launches"]
        e["traffic_bytes_per_launch"] = (e["dram_bytes_read"] + e["dram_bytes_write"]) / n
        e["gpu_time_ms_under_ncu"] /= n
        e["dram_bytes_read"] /= n
        e["dram_bytes_write"] /= n
    json.dump({"command": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                          "-k regex:k_fir|k_trace|k_chain -c 10 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e",
               "workload": "C5 2048x2048x4096, 8 bands, 1 x B200, final round-2 build", "cube": [2048, 2048, 4096],
               "kernels": kernels}, open(os.path.join(P, dst), "w"), indent=1)
    shutil.copy(path, os.path.join(P, dst.replace("_traffic.json", "_ncu_cube_kernels_c5.csv")))
    for k, e in kernels.items():
        print(f"{k}: {e['traffic_bytes_per_launch'] / 1e9:.2f} GB per launch, {e['gpu_time_ms_under_ncu']:.1f} ms under ncu")


def slab_trace(prefix, n, dst):
    files = sorted(glob.glob(os.path.join(G, f"{prefix}.rank*.csv")))
    if not files:
        return
    out = [f"# THZ_SLAB_TRACE of config 5 on {n} GPUs: per-launch time stamps of band 0 (47x57 PSF, 423 iterations = 846 launches),",
           "# medians over launch ranges; period = start-to-start, kernel = first CTA start to last CTA end, wait = longest halo wait",
           "# of a boundary CTA, bdone = last boundary CTA (rows pushed, version published) after kernel start; microseconds"]
    for f in files:
        rows = list(csv.DictReader(open(f)))
        r = os.path.basename(f).split("rank")[1].split(".")[0]
        per = [int(rows[i + 1]["start_ns"]) - int(rows[i]["start_ns"]) for i in range(len(rows) - 1)]
        dur = [int(x["end_ns"]) - int(x["start_ns"]) for x in rows]
        wait = [int(x["halo_wait_ns"]) for x in rows]
        bd = [int(x["boundary_end_ns"]) - int(x["start_ns"]) for x in rows]
        parts = []
        for lo, hi in ((4, 26), (30, 90), (100, 250), (260, 500), (520, 840)):
            q = lambda v: st.median(v[lo:hi]) / 1e3
            parts.append(f"[{lo}-{hi}] period {q(per):.1f} kernel {q(dur):.1f} wait {q(wait):.1f} bdone {q(bd):.1f}")
        out.append(f"rank {r}: " + " | ".join(parts))
    open(os.path.join(P, dst), "w").write("\n".join(out) + "\n")
    print(dst)


def ncu_summary(rep, dst, title):
    path = os.path.join(G, rep)
    if os.path.exists(path):
        subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), path, os.path.join(P, dst), title], check=False)
        print(dst)


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "z"
    copy_bench(f"{tag}_bench_c5.json", "r02_bench_c5_1gpu.json")
    copy_bench(f"{tag}_bench_reference_arm.json", "r02_bench_reference_arm.json")
    for c in ("c1", "c2", "c3", "c4"):
        copy_bench(f"{tag}_bench_{c}.json", f"r02_bench_{c}_1gpu.json")
    traffic_from_csv(f"{tag}_ncu_cube_kernels_c5.csv", "r02_traffic.json")
    if os.path.exists(os.path.join(G, f"{tag}_ncu_launches_512.csv")):
        shutil.copy(os.path.join(G, f"{tag}_ncu_launches_512.csv"), os.path.join(P, "r02_ncu_launches_full_step_512x512.csv"))
    ncu_summary(f"{tag}_prof_cube512.ncu-rep", "r02_ncu_full_cube_kernels_512x512.txt",
                "ncu --set full --clock-control none --import-source on, cube kernels of the final round-2 build on a "
                "512 x 512 x 4096 cube (1/16 of config 5, same work per trace)")
    ncu_summary(f"{tag}_prof_trace_c4.ncu-rep", "r02_ncu_full_k_trace_fused_c4.txt",
                "ncu --set full --clock-control none --import-source on -k regex:k_trace_fused, config 4 (1024 x 1024 x 2048, "
                "default chain, no deconvolution)")
    ncu_summary(f"{tag}_prof_rl.ncu-rep", "r02_ncu_full_k_rl_multi.txt",
                "ncu --set full --clock-control none --import-source on -k regex:k_rl_multi -s 600 -c 2, config 5 (band 0 alone, 47x57 PSF)")
