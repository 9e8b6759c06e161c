"""profiles/ncu_traffic.json from ncu launch lists of `bench.py --one-step --method M`.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \\
        --log-file gpurun_out/traffic_M.csv python bench.py --one-step --method M
    python scripts/ncu_traffic.py gpurun_out/traffic_*.csv        # writes profiles/ncu_traffic.json

Per method and bench span (the tags of bench.py's roofline): DRAM bytes read + written by the vlmc kernels of ONE step.
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def tag_of(kernel):
    k = kernel
    if "colstats_kernel" in k or "colstats_batch_kernel" in k:
        m = re.search(r"colstats_(?:batch_)?kernel<[^,]+,\s*\(?(?:bool\))?\s*(\w+)", k)
        d = m.group(1) if m else "0"
        return "dsnot_stats" if d in ("1", "true") else "sqnorm_accum"
    if "hessian_syrk" in k:
        return "hessian_accum"
    if any(x in k for x in ("nm_batch_kernel", "nm_kernel", "rowselect", "sqrt_vec", "mean_finalize", "thr_")):
        return "wanda_select"
    if "dsnot_" in k:
        return "dsnot_refine"
    if any(x in k for x in ("potrf", "gemm3x", "flip_", "obs_", "diag_prepare", "add_diag", "lower_to_upper")):
        return "sparsegpt_chains"
    return None


def main(paths):
    out = {"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum on `bench.py --one-step --method M` (1 GPU, "
                     "calib_batch 128); bytes of ONE step per bench span", "per_step_bytes": {}, "per_step_ms_under_ncu": {},
           "kernels": {}}
    for path in paths:
        method = re.sub(r"^.*traffic_|\.csv$", "", os.path.basename(path))
        rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
        hdr = rows[0]
        ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        byt = collections.defaultdict(float)
        ms = collections.defaultdict(float)
        per_kernel = collections.defaultdict(lambda: [0, 0.0, 0.0])
        for r in rows[1:]:
            name = r[ki]
            if "vlmc::" not in name:
                continue
            tag = tag_of(name)
            if tag is None:
                continue
            v = float(r[vi].replace(",", ""))
            unit = r[ui]
            short = re.sub(r"\(.*$", "", name.replace("void ", "").replace("vlmc::", ""))[:60]
            if r[mi].startswith("dram__bytes"):
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
                byt[tag] += v * scale
                per_kernel[short][1] += v * scale
            elif r[mi].startswith("gpu__time_duration"):
                scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
                ms[tag] += v * scale
                per_kernel[short][0] += 1
                per_kernel[short][2] += v * scale
        out["per_step_bytes"][method] = dict(byt)
        out["per_step_ms_under_ncu"][method] = dict(ms)
        out["kernels"][method] = {k: {"launches": v[0], "dram_bytes": v[1], "ms_under_ncu": v[2]} for k, v in per_kernel.items()}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out["per_step_bytes"], indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
