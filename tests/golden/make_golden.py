#!/usr/bin/env python
"""Generate tests/golden/ from the reference's own test artefacts (run in the build container only).

/root/reference does not exist on the GPU box, so every golden vector, known-answer value and published count
that the tests use is extracted here ONCE and committed:

  * test/reftest-files/test_1.bson                    -> part1_test_1.json      (test/part1.jl:24-40, atol 1e-5)
  * benchmark-results/error_vs_*_results.csv          -> part1_kats.json        (17-digit point values)
  * benchmark-results/bench_diffusion_scaling_*.csv   -> part1_kats.json        (timed iteration counts, decoded
        as Work / (ranks*27*(nx-2)(ny-2)(nz-2)), scripts-part1/part1_kernel_programming.jl:210)
  * benchmark-results/bench_multigrid_*.csv           -> part2_published.json   (published solve times)
  * test/reftest-files/fortran/*.bin                  -> fortran/*.bin          (data files, byte-for-byte copies;
        format = scripts-part2/part2_utils.jl:11-19: Int32 nx, Int32 ny, nx*ny Float64 column-major)

Only data is copied, never reference source code.
"""
import csv
import json
import os
import shutil
import struct
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def bson_doc(buf, off=0):
    """Minimal BSON decoder (documents, arrays, strings, binary, int32/int64, double, bool, null)."""
    (n,) = struct.unpack_from("<i", buf, off)
    end = off + n - 1
    p = off + 4
    out = {}
    while p < end:
        t = buf[p]
        p += 1
        z = buf.index(b"\x00", p)
        key = buf[p:z].decode()
        p = z + 1
        if t == 0x01:
            (v,) = struct.unpack_from("<d", buf, p); p += 8
        elif t == 0x02:
            (ln,) = struct.unpack_from("<i", buf, p); v = buf[p + 4:p + 4 + ln - 1].decode(); p += 4 + ln
        elif t in (0x03, 0x04):
            v, p2 = bson_doc(buf, p)
            if t == 0x04:
                v = [v[k] for k in sorted(v, key=int)]
            p = p2
        elif t == 0x05:
            (ln,) = struct.unpack_from("<i", buf, p); v = bytes(buf[p + 5:p + 5 + ln]); p += 5 + ln
        elif t == 0x08:
            v = bool(buf[p]); p += 1
        elif t == 0x0A:
            v = None
        elif t == 0x10:
            (v,) = struct.unpack_from("<i", buf, p); p += 4
        elif t == 0x12:
            (v,) = struct.unpack_from("<q", buf, p); p += 8
        else:
            raise ValueError(f"BSON type {t:#x} not handled")
        out[key] = v
    return out, off + n


def julia_array(d):
    assert d["tag"] == "array" and d["type"]["name"] == ["Core", "Float64"]
    size = d["size"]
    n = 1
    for s in size:
        n *= s
    vals = list(struct.unpack(f"<{n}d", d["data"]))
    return {"size": size, "data_column_major": vals}


def main():
    os.makedirs(os.path.join(OUT, "fortran"), exist_ok=True)

    # ---- test_1.bson ------------------------------------------------------------------------------------
    buf = open(os.path.join(REF, "test/reftest-files/test_1.bson"), "rb").read()
    doc, _ = bson_doc(buf)
    g = {"source": "test/reftest-files/test_1.bson", "atol": 1e-5,
         "inds_1based": [1, 4, 7, 10, 13, 16, 18, 21, 24, 27, 30, 32], "z_index_1based": 15,
         "X": julia_array(doc["X"]), "H": julia_array(doc["H"])}
    json.dump(g, open(os.path.join(OUT, "part1_test_1.json"), "w"), indent=1)

    # ---- point-value KATs + published iteration counts ------------------------------------------------------
    kats = {"point_values_vs_grid_size": [], "point_values_vs_tolerance": [], "scaling_counts": []}
    with open(os.path.join(REF, "benchmark-results/error_vs_grid_size_experiment_results.csv")) as f:
        for row in csv.DictReader(f):
            kats["point_values_vs_grid_size"].append({"nx": int(row["nx"]), "val": float(row["val"]),
                                                      "val_str": row["val"], "ttot": 2.0, "tol": 1e-6})
    with open(os.path.join(REF, "benchmark-results/error_vs_tolerance_experiment_results.csv")) as f:
        for row in csv.DictReader(f):
            kats["point_values_vs_tolerance"].append({"nx": 128, "tol": float(row["tol"]), "val": float(row["val"]),
                                                      "val_str": row["val"], "ttot": 2.0})
    strong_dims = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}
    seen = set()
    for dev in ("gpu", "cpu"):
        with open(os.path.join(REF, f"benchmark-results/bench_diffusion_scaling_{dev}.csv")) as f:
            for lineno, row in enumerate(csv.DictReader(f), start=2):
                ranks = int(row["n_mpi_ranks"])
                strong = row["strong_scaling"] == "true"
                dims = strong_dims[ranks]
                local = tuple(128 // d for d in dims) if strong else (128, 128, 128)
                cells = (local[0] - 2) * (local[1] - 2) * (local[2] - 2)
                iters = float(row["Work"]) / (ranks * 27 * cells)
                assert abs(iters - round(iters)) < 1e-6, (dev, lineno, iters)
                key = (ranks, strong)
                rec = {"ranks": ranks, "dims": list(dims), "local": list(local), "strong_scaling": strong,
                       "timed_iters": int(round(iters)), "ttot": 2.0, "tol": 1e-6, "warmup_steps": 3,
                       "source": f"benchmark-results/bench_diffusion_scaling_{dev}.csv:{lineno}",
                       "delta_t_s": float(row["delta_t"]), "use_shared_memory": row["use_shared_memory"] == "true",
                       "use_gpu": dev == "gpu"}
                if key in seen:
                    prev = [k for k in kats["scaling_counts"] if (k["ranks"], k["strong_scaling"]) == key][0]
                    assert prev["timed_iters"] == rec["timed_iters"], (prev, rec)
                    prev.setdefault("also", []).append({k: rec[k] for k in
                                                        ("source", "delta_t_s", "use_shared_memory", "use_gpu")})
                else:
                    seen.add(key)
                    kats["scaling_counts"].append(rec)
    json.dump(kats, open(os.path.join(OUT, "part1_kats.json"), "w"), indent=1)

    # ---- published multigrid timings -----------------------------------------------------------------------
    pub = []
    for name in ("bench_multigrid_gpu.csv", "bench_multigrid_gpu_V100.csv", "bench_multigrid_cpu.csv"):
        with open(os.path.join(REF, "benchmark-results", name)) as f:
            for lineno, row in enumerate(csv.DictReader(f), start=2):
                if int(row["l"]) == 2 and int(row["k"]) >= 10:
                    pub.append({"file": name, "line": lineno, **{k: row[k] for k in row}})
    json.dump(pub, open(os.path.join(OUT, "part2_published.json"), "w"), indent=1)

    # ---- Fortran goldens (data) -----------------------------------------------------------------------------
    src = os.path.join(REF, "test/reftest-files/fortran")
    for name in sorted(os.listdir(src)):
        shutil.copyfile(os.path.join(src, name), os.path.join(OUT, "fortran", name))
        os.chmod(os.path.join(OUT, "fortran", name), 0o644)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
