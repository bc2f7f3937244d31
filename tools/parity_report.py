#!/usr/bin/env python
"""Turns the PYSDR_PARITY_LOG written by tests/util.assert_parity during `pytest -m gpu` into a per-test table of the
worst margins (max-abs relative error vs its gate, difference SNR vs its gate).  Usage: parity_report.py LOG.jsonl > report.txt"""
import collections
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1])]
per = collections.OrderedDict()
for r in rows:
    t = r["test"].split("[")[0]
    e = per.setdefault(t, {"n": 0, "rel": 0.0, "rel_tol": 0.0, "snr": float("inf"), "snr_min": 0.0})
    e["n"] += 1
    if r["rel"] >= e["rel"]:
        e["rel"], e["rel_tol"] = r["rel"], r["rel_tol"]
    if r["snr_db"] is not None and r["snr_db"] < e["snr"]:
        e["snr"], e["snr_min"] = r["snr_db"], r["snr_min"]
gates = collections.Counter((r["rel_tol"], r["snr_min"]) for r in rows)
print("%d comparisons; gates used (rel_tol, snr_min dB): %s" % (len(rows), dict(gates)))
print("%-78s %5s %11s %8s %9s %6s" % ("test", "n", "worst rel", "gate", "min SNR", "gate"))
for t, e in per.items():
    print("%-78s %5d %11.3e %8.0e %9.1f %6.0f" % (t, e["n"], e["rel"], e["rel_tol"], e["snr"], e["snr_min"]))
