"""Phase view of a kernel timeline written by tools/timeline.py (<out>_kernels.csv): wall time of the step's phases, delimited by
marker kernels, and inside every phase the time during which NO kernel runs / only narrow-class kernels run."""
import csv
import sys

rows = []
for r in csv.DictReader(open(sys.argv[1])):
    rows.append((float(r["start_us"]), float(r["dur_us"]), int(r["stream"]), r["name"].strip('"')))
rows.sort()


def first(name, after=0.0):
    for s, d, st, n in rows:
        if name in n and s >= after:
            return s
    return None


def last_end(name):
    return max((s + d for s, d, st, n in rows if name in n), default=None)


t0 = first("img_to_hwc4") or first("stem_pool") or first("stem_conv")   # first kernel of the conv backbone (fused / two-kernel stem)
marks = [("forward: conv backbone (+BERT beside it)", t0, first("groupnorm_tokens_fwd")),
         ("forward: encoder + decoder + heads", first("groupnorm_tokens_fwd"), first("box_loss") or first("attn_small_bwd")),
         ("backward: heads + decoder + encoder", first("box_loss") or first("attn_small_bwd"), first("groupnorm_tokens_bwd")),
         ("backward: conv backbone (+BERT beside it)", first("groupnorm_tokens_bwd"), first("scale_copy_check")),
         ("hand-over", first("scale_copy_check"), last_end("scale_copy_check"))]
# (single-process timelines: under data parallelism the hand-over is sliced and overlaps the backward, the last two rows then overlap)
for name, a, b in marks:
    if a is None or b is None:
        print(f"{name}: markers missing")
        continue
    ks = [(s, d, n) for s, d, st, n in rows if s + d > a and s < b]
    # coverage: union of intervals
    ev = sorted((max(s, a), min(s + d, b)) for s, d, n in ks)
    busy, cur_s, cur_e = 0.0, None, None
    for s, e in ev:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        busy += cur_e - cur_s
    tot = sum(min(s + d, b) - max(s, a) for s, d, n in ks)
    print(f"{name:48s} {b - a:8.1f} us   kernels {len(ks):4d}   some kernel running {busy:8.1f} us   sum of durations {tot:8.1f} us")
if marks[-1][2] is not None and t0 is not None:
    print(f"step (first backbone kernel .. hand-over end): {marks[-1][2] - t0:.1f} us")
