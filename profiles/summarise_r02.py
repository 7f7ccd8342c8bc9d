"""Turns the raw artefacts of profiles/collect_r02.sh (gpurun_out/r02_*) into the committed profiles/r02_* files:
bench JSON lines (pretty-printed), launch-list summary, ncu text summaries and the per-launch DRAM traffic files
bench.py reads for roofline.traffic."""
import csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def last_json(path):
    if not os.path.exists(path):
        return None
    lines = [l for l in open(path).read().strip().split("\n") if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None

for name in ("r02_bench_c3", "r02_bench_c1", "r02_bench_c2", "r02_bench_c5", "r02_bench_c4", "r02_bench_reference_arm",
             "r02_bench_c3_2gpu", "r02_bench_c3_4gpu", "r02_bench_c3_8gpu", "r02_bench_c5_2gpu", "r02_bench_c5_4gpu", "r02_bench_c5_8gpu"):
    d = last_json(os.path.join(G, name + ".json"))
    if d is None:
        continue
    if name == "r02_bench_c4":
        json.dump(dict(benchmark="C4 kNN sweep (bench.py --workload c4)", rows=d.pop("sweep")), open(os.path.join(P, "r02_knn_sweep.json"), "w"), indent=1)
    json.dump(d, open(os.path.join(P, name + ".json"), "w"), indent=1)
    print(name, "value %.1f" % d["value"], "e2e %.1f" % d["e2e"]["value"])

d = os.path.join(G, "r02_reg_compare_c3.json")
if os.path.exists(d):
    shutil.copy(d, os.path.join(P, "r02_reg_compare_c3.json"))
log = os.path.join(G, "r02_pytest_gpu.log")
if os.path.exists(log):
    open(os.path.join(P, "r02_pytest_gpu_tail.txt"), "w").write("".join(open(log).readlines()[-4:]))

csvp = os.path.join(G, "r02_launches_bench_c3.csv")
if os.path.exists(csvp):
    # keep what follows the last cache-fill kernel (the keyframe cache is filled once per session, by the first call)
    lines = open(csvp).read().split("\n")
    hdr_i = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    body = [l for l in lines[hdr_i + 1:] if l.strip()]
    last_fill = max([i for i, l in enumerate(body) if "gather_tf_kernel" in l or "bbox_tf_kernel" in l] + [-1])
    steady = os.path.join(P, "r02_launches_bench_c3.csv")
    # ... and drop the radix-sort micro-benchmark bench.py runs after the timed steps (roofline_sort_pass: 8192-pair tiles,
    # which no step of this workload launches any more)
    keep = [l for l in body[last_fill + 1:] if "rs_onesweep_kernel<512>" not in l and "bench_fill_kernel" not in l and
            "rs_global_hist_kernel" not in l]
    open(steady, "w").write("\n".join(lines[:hdr_i + 1] + keep) + "\n")
    body = body[:last_fill + 1] + keep
    out = subprocess.run([sys.executable, os.path.join(P, "launch_summary.py"), steady], capture_output=True, text=True).stdout
    head = ("launch list summary of `bench.py --steps 2 --warmup 1 --no-cpu-baseline --sustain-seconds 0` (ncu --metrics "
            "gpu__time_duration.sum --clock-control none; the %d launches after the one-time keyframe-cache fill of %d launches; "
            "cold-cache, serialised: compare SHARES, not absolutes)\n" % (len(body) - last_fill - 1, last_fill + 1))
    open(os.path.join(P, "r02_launches_bench_c3_summary.txt"), "w").write(head + out)
    print(out[:600])

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))

def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

for rep, txt, traffic, pairs_key in (("r02_ncu_register.ncu-rep", "r02_ncu_register_summary.txt", "r02_register_kernel_traffic.json", "queries"),
                                     ("r02_ncu_bucket.ncu-rep", "r02_ncu_bucket_summary.txt", "r02_bucket_kernel_traffic.json", "points"),
                                     ("r02_ncu_sort.ncu-rep", "r02_ncu_sort_summary.txt", "r02_sort_kernel_traffic.json", "pairs")):
    rp = os.path.join(G, rep)
    if not os.path.exists(rp):
        continue
    out = subprocess.run([sys.executable, os.path.join(P, "ncu_summary.py"), rp], capture_output=True, text=True).stdout
    open(os.path.join(P, txt), "w").write(out)
    vals, units = raw(rp)
    rd = to_bytes(vals["dram__bytes_read.sum"], units["dram__bytes_read.sum"])
    wr = to_bytes(vals["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
    b3 = last_json(os.path.join(G, "r02_bench_c3.json")) or {}
    n = None
    import re
    if pairs_key == "pairs":
        m = re.search(r"over (\d+) pairs", (b3.get("roofline_sort_pass") or {}).get("kernel", ""))
        n = int(m.group(1)) if m else None
    elif pairs_key == "points":
        m = re.search(r"(\d+) points in", (b3.get("roofline_map_build_kernel") or {}).get("kernel", ""))
        n = int(m.group(1)) if m else None
    else:
        n = (b3.get("workload_stats") or {}).get("queries_per_scan")
    json.dump(dict(kernel=vals.get("Kernel Name", ""), pairs=n, dram_bytes_read=rd, dram_bytes_write=wr, dram_bytes_per_launch=rd + wr,
                   duration_us=vals.get("gpu__time_duration.sum"), source="ncu --set full --clock-control none, one launch (" + rep + ")"),
              open(os.path.join(P, traffic), "w"), indent=1)
    print(txt, "dram R %.1f MB W %.1f MB" % (rd / 1e6, wr / 1e6))

# the bench line was printed before these captures existed on the box: fill its `traffic` fields from them
b3p = os.path.join(P, "r02_bench_c3.json")
if os.path.exists(b3p):
    b3 = json.load(open(b3p))
    for key, tf, nkey in (("roofline", "r02_register_kernel_traffic.json", "queries"), ("roofline_map_build_kernel", "r02_bucket_kernel_traffic.json", "points"),
                          ("roofline_sort_pass", "r02_sort_kernel_traffic.json", "pairs")):
        tp = os.path.join(P, tf)
        if key in b3 and b3[key] and os.path.exists(tp):
            t = json.load(open(tp))
            b3[key]["traffic"] = t.get("dram_bytes_per_launch")
            b3[key]["traffic_source"] = "profiles/" + tf
    json.dump(b3, open(b3p, "w"), indent=1)
