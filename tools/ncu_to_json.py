"""profiles/ncu_summary.json for bench.py (bench.ncu_summary): the profiler-derived figures of the bench line, taken from ncu CSV
captures committed under profiles/ -- never from constants in the source.

    python tools/ncu_to_json.py --gemm profiles/rNN_gemm_dram.csv --attn profiles/rNN_ncu_attention.csv [--lib reftr_b200/libreftr_b200.so]

--gemm : `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,... -k regex:umma_gemm_kernel|gemm_skinny --csv` of ONE step
--attn : tools/ncu_summary.py output of an `ncu --set full` capture that contains the encoder attention kernel(s)"""
import argparse
import csv
import hashlib
import json
import os
import re


def load_long(path):
    """ncu --csv long format (one row per kernel x metric) -> [{kernel, id, metrics{name: (value, unit)}}]"""
    lines = [l for l in open(path) if l.startswith('"')]
    out = {}
    for row in csv.DictReader(lines):
        k = row["ID"]
        d = out.setdefault(k, {"kernel": re.sub(r"\(.*", "", row["Kernel Name"]), "metrics": {}})
        try:
            d["metrics"][row["Metric Name"]] = (float(row["Metric Value"].replace(",", "")), row["Metric Unit"])
        except ValueError:
            pass
    return list(out.values())


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v, unit):
    return v * {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gemm")
    ap.add_argument("--attn")
    ap.add_argument("--lib", default="reftr_b200/libreftr_b200.so")
    ap.add_argument("--out", default="profiles/ncu_summary.json")
    a = ap.parse_args()
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from reftr_b200._lib import kernel_source_hash
    res = {"kernel_src_sha256": kernel_source_hash()}  # the sources the profiled library was built from (run this right after the capture)
    if a.gemm:
        ks = load_long(a.gemm)
        n = len(ks)
        rd = sum(to_bytes(*k["metrics"]["dram__bytes_read.sum"]) for k in ks if "dram__bytes_read.sum" in k["metrics"])
        wr = sum(to_bytes(*k["metrics"]["dram__bytes_write.sum"]) for k in ks if "dram__bytes_write.sum" in k["metrics"])
        us = sum(to_us(*k["metrics"]["gpu__time_duration.sum"]) for k in ks if "gpu__time_duration.sum" in k["metrics"])
        res["gemm_family"] = {"launches": n, "dram_bytes_per_launch": (rd + wr) / max(n, 1), "dram_read_gb_per_step": rd / 1e9, "dram_write_gb_per_step": wr / 1e9,
                              "sum_duration_us_under_ncu": us, "source": a.gemm,
                              "note": f"(dram__bytes_read.sum + dram__bytes_write.sum) over the {n} GEMM-family launches of one cfg2 train-mode step / {n}; "
                                      f"{(rd + wr) / 1e9:.2f} GB per step (reads {rd / 1e9:.2f} GB); ncu capture {a.gemm}"}
    if a.attn:
        rows = list(csv.reader(open(a.attn)))
        hdr = rows[0]
        col = {h: i for i, h in enumerate(hdr)}
        enc = [r for r in rows[2:] if "attn_fwd_tc" in r[col["Kernel Name"]] or "enc_mha" in r[col["Kernel Name"]]]
        if enc:
            r = enc[0]

            def f(name):
                try:
                    return float(r[col[name]].replace(",", ""))
                except Exception:
                    return None
            res["encoder_mha"] = {"kernel": r[col["Kernel Name"]][:80], "grid": r[col["Grid Size"]], "duration_us": f("gpu__time_duration.sum"),
                                  "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                                  "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                                  "dram_pct_of_peak": f("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                                  "registers_per_thread": f("launch__registers_per_thread"), "ncu_csv": a.attn,
                                  "config": "cfg2 train mode: S=420, head_dim 32, B*H=128, dropout on P"}
    json.dump(res, open(a.out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
