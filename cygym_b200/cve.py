"""Synthetic CVE table with CVE.csv's schema.

The reference reads `CVE.csv` (CDSimulator.py:36), a Kaggle/NVD export produced by
parse_json.py:33-49 with 15 columns, of which the simulator reads `matchCriteriaId`,
`exploitabilityScore` and `impactScore` (CDSimulator.py:499-503, :508-518, :561-566, :578-590).
That file is not available offline, so networks are built from a seeded table of the same shape.
The two ids hard-coded at volt_typhoon_env.py:22-23 are always rows 0 and 1.
"""
import csv
import uuid

import numpy as np

VOLT_CVE_ID = "ED3A999C-9184-4D27-A62E-3D8A3F0D4F27"
VOLT_DC_CVE_ID = "0A5713AE-B7C5-4599-8E4F-9C235E73E5F6"

COLUMNS = [
    "CVE_id", "source_identifier", "published_time", "lastModified_time", "baseScore",
    "baseSeverity", "exploitabilityScore", "impactScore", "matchCriteriaId",
    "versionStartIncluding", "versionEndExcluding", "type", "vendor", "product", "version",
]


def synthetic_cve_table(n_rows=256, seed=7):
    """Returns a dict column -> list with the 15 columns of parse_json.py:33-49."""
    rng = np.random.default_rng(seed)
    t = {c: [] for c in COLUMNS}
    for i in range(n_rows):
        expl = round(float(rng.uniform(0.5, 3.9)), 1)
        base = round(min(10.0, expl + float(rng.uniform(1.0, 5.0))), 1)
        if i == 0:
            mcid = VOLT_CVE_ID
        elif i == 1:
            mcid = VOLT_DC_CVE_ID
        else:
            mcid = str(uuid.UUID(int=int.from_bytes(rng.bytes(16), "little"))).upper()
        t["CVE_id"].append(f"CVE-2024-{10000 + i}")
        t["source_identifier"].append("synthetic@cygym-b200")
        t["published_time"].append("2024-01-01T00:00:00.000")
        t["lastModified_time"].append("2024-01-02T00:00:00.000")
        t["baseScore"].append(base)
        t["baseSeverity"].append("HIGH" if base >= 7 else "MEDIUM")
        t["exploitabilityScore"].append(expl)
        t["impactScore"].append(round(float(rng.uniform(1.4, 5.9)), 1))
        t["matchCriteriaId"].append(mcid)
        t["versionStartIncluding"].append("")
        t["versionEndExcluding"].append("")
        t["type"].append("a")
        t["vendor"].append("synthetic")
        t["product"].append(f"product_{i}")
        t["version"].append("*")
    return t


def write_csv(table, path):
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(COLUMNS)
        for i in range(len(table["CVE_id"])):
            w.writerow([table[c][i] for c in COLUMNS])
    return path


def read_csv(path):
    with open(path, newline="") as f:
        r = csv.DictReader(f)
        t = {c: [] for c in COLUMNS}
        for row in r:
            for c in COLUMNS:
                v = row[c]
                t[c].append(float(v) if c in ("baseScore", "exploitabilityScore", "impactScore") else v)
    return t
