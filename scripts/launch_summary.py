"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list of a bench.py run into what profiles/ keeps:

  <out>.csv   the launches of ONE forward (from a text_embed kernel to the launch before the next one), as ncu printed them
  <out>.json  launches / total ns / share per kernel family over ALL complete forwards found in the list

usage: python scripts/launch_summary.py gpurun_out/r2/launches.csv profiles/r2_launches
"""
import csv
import json
import re
import sys


def family(name):
    for key in ("gemm_persistent_kernel", "gemm_pair_kernel", "gemm_chain_kernel", "ln_residual_kernel", "self_attention_kernel",
                "co_attention_kernel", "attention_f32_kernel", "text_embed_kernel", "image_pack_kernel", "region_pack_kernel", "rowdot_kernel",
                "gather_state_kernel"):
        if key in name:
            return key
    return re.sub(r"<.*", "", name).split("::")[-1][:60]


def main(src, out):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    starts = [i for i, r in enumerate(rows) if "text_embed_kernel" in r["Kernel Name"]]
    if len(starts) < 2:
        raise SystemExit("no complete forward in the list")
    # forwards = runs between consecutive text_embed launches with the same length as the most common one
    lens = [b - a for a, b in zip(starts, starts[1:])]
    common = max(set(lens), key=lens.count)
    fwd = [(a, b) for a, b in zip(starts, starts[1:]) if b - a == common]
    fam = {}
    for a, b in fwd:
        for r in rows[a:b]:
            f = fam.setdefault(family(r["Kernel Name"]), [0, 0.0])
            f[0] += 1
            f[1] += float(r["Metric Value"])
    total = sum(v[1] for v in fam.values())
    summary = {"source": src, "forwards": len(fwd), "launches_per_forward": common, "ns_per_forward": total / len(fwd),
               "families": {k: {"launches_per_forward": v[0] / len(fwd), "ns_per_forward": v[1] / len(fwd), "share": v[1] / total}
                            for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1])}}
    json.dump(summary, open(out + ".json", "w"), indent=1)
    a, b = fwd[-1]
    with open(out + ".csv", "w") as f:
        f.write(lines[0])
        for l in lines[1 + a:1 + b]:
            f.write(l)
    print(json.dumps(summary))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
