"""Developer tool: AEC_PARITY_LOG (one JSON row per float-net comparison of tests/parity.py) + the bench line's parity_check
-> profiles/parity_r2.json.   usage: make_parity_json.py parity.jsonl pytest.txt bench.json out.json "<what ran>" """
import json
import sys


def main(jsonl, pytest_txt, bench_json, out, what):
    rows = [json.loads(l) for l in open(jsonl) if l.strip()]
    tail = [l.strip() for l in open(pytest_txt) if " passed" in l or " failed" in l]
    bench = json.loads(open(bench_json).read().strip().splitlines()[-1])
    bad = [r["case"] for r in rows if any(v for k, v in r.items() if k.endswith("_unexplained"))]
    doc = {"source": "AEC_PARITY_LOG of `python -m pytest tests -m gpu` on one B200 (tools/final_capture.sh, %s): %s; one row per float-net "
                     "comparison of tests/parity.py, *_unexplained must be 0, *_roots are near ties of the oracle's own values.  "
                     "bench_parity_check = the parity_check object of the bench line of the same call" % (what, tail[-1] if tail else "?"),
           "rows": rows, "cases_with_unexplained_mismatches": bad, "bench_parity_check": bench.get("parity_check")}
    json.dump(doc, open(out, "w"), indent=1)
    print(len(rows), "rows,", len(bad), "with unexplained mismatches;", tail[-1] if tail else "")


if __name__ == "__main__":
    main(*sys.argv[1:6])
