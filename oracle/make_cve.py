"""Synthetic CVE.csv with the reference's schema (TEST INFRASTRUCTURE ONLY).

Columns are the 15 that parse_json.py:33-49 writes.  Only matchCriteriaId,
exploitabilityScore and impactScore are read by the simulator
(CDSimulator.py:499-503, :508-518, :561-566, :578-590).  The two ids hard-coded
at volt_typhoon_env.py:22-23 must be present or target-mode exploit generation
returns nothing (CDSimulator.py:579-581).  The real Kaggle/NVD file is not
available offline, so scores are drawn from a seeded generator.
"""
import csv
import random
import uuid

VOLT_CVE_ID = "ED3A999C-9184-4D27-A62E-3D8A3F0D4F27"
VOLT_DC_CVE_ID = "0A5713AE-B7C5-4599-8E4F-9C235E73E5F6"

COLUMNS = [
    "CVE_id", "source_identifier", "published_time", "lastModified_time", "baseScore",
    "baseSeverity", "exploitabilityScore", "impactScore", "matchCriteriaId",
    "versionStartIncluding", "versionEndExcluding", "type", "vendor", "product", "version",
]


def write_cve_csv(path, n_rows=64, seed=7, volt_score=3.9, volt_dc_score=3.9):
    rng = random.Random(seed)
    rows = []

    def row(i, mcid, expl):
        base = round(min(10.0, expl + rng.uniform(1.0, 5.0)), 1)
        return {
            "CVE_id": f"CVE-2024-{10000 + i}",
            "source_identifier": "synthetic@cygym-b200",
            "published_time": "2024-01-01T00:00:00.000",
            "lastModified_time": "2024-01-02T00:00:00.000",
            "baseScore": base,
            "baseSeverity": "HIGH" if base >= 7 else "MEDIUM",
            "exploitabilityScore": expl,
            "impactScore": round(rng.uniform(1.4, 5.9), 1),
            "matchCriteriaId": mcid,
            "versionStartIncluding": "",
            "versionEndExcluding": "",
            "type": "a",
            "vendor": "synthetic",
            "product": f"product_{i}",
            "version": "*",
        }

    rows.append(row(0, VOLT_CVE_ID, volt_score))
    rows.append(row(1, VOLT_DC_CVE_ID, volt_dc_score))
    for i in range(2, n_rows):
        mcid = str(uuid.UUID(int=rng.getrandbits(128))).upper()
        rows.append(row(i, mcid, round(rng.uniform(0.5, 3.9), 1)))
    with open(path, "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=COLUMNS)
        w.writeheader()
        for r in rows:
            w.writerow(r)
    return path
