#!/usr/bin/env python
"""profiles/<tag>_traffic.json from `ncu -i X.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,
gpu__time_duration.sum,...` dumps (one per workload): DRAM bytes per launch of the kernels bench.py's roofline names.

    python profiles/make_traffic.py r02 T=profiles/r02_ncu_full_T.csv [cfg2=...]
"""
import csv, json, os, re, sys

PHASE_OF = [("k_encoder_forward", "encoder_forward"), ("k_score", "score"), ("k_tc_dq", "contract_dq"), ("k_tc_dc2", "contract_dc"),
            ("k_tc_dc2", "contract_dc"), ("k_rows_chunk<0", "w_update"), ("k_rows_chunk<(int)0", "w_update"), ("k_rows_chunk<1", "entity_update"),
            ("k_rows_chunk<(int)1", "entity_update"), ("k_tc_bwd_finish", "backward_finish"), ("k_dense_finalize", "dense_finalize"),
            ("k_dense_apply", "dense_apply")]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
TAG = sys.argv[1]
for arg in sys.argv[2:]:
    wl, path = arg.split("=")
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen_bil = 0
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        phase = None
        if "k_tc_bilinear" in name:
            phase = "contract_forward" if seen_bil % 2 == 0 else "contract_recompute"
            seen_bil += 1
        else:
            for key, ph in PHASE_OF:
                if key in name:
                    phase = ph
                    break
        if phase is None:
            continue
        rd = float(r[idx["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[idx["dram__bytes_write.sum"]]]
        e = out.setdefault(wl, {}).setdefault(phase, {"kernel": re.sub(r"\(.*", "", name).replace("void ", "").strip(), "launches": 0, "sum": 0.0})
        e["launches"] += 1
        e["sum"] += rd + wr
for wl in out:
    for ph, e in out[wl].items():
        e["dram_bytes_per_launch"] = e.pop("sum") / e["launches"]
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), TAG + "_traffic.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
