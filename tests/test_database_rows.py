"""include/dagsfm_b200/database_rows.hpp: the batched output stage of SiftFeatureMatcher::Match.  The rows it builds
are inserted into a SQLite database created with the reference's schema and statements
(src/base/database.cc:1233-1261, 1121-1130) and read back the way Database::ReadMatches (:455-473) and
Database::ReadTwoViewGeometry (:494-533) do -- including the column swap / TwoViewGeometry::Invert for pairs stored
in swapped order -- and must return what went in.  Expected rows are also built by an independent Python
restatement of Database::WriteMatches / WriteTwoViewGeometry (:680-755)."""
import sqlite3
import struct
import subprocess
from pathlib import Path

import numpy as np

from dagsfm_b200.verification import POSE_DTYPE, RESULT_DTYPE

ROOT = Path(__file__).resolve().parent.parent
KMAX = 2147483647


def _exe(tmp: Path) -> Path:
    exe = tmp / "database_rows_test"
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests/cpp/database_rows_test.cc"),
                        "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def quat_R(q):
    n = np.linalg.norm(q)
    q = np.array([1.0, q[1], q[2], q[3]]) if n == 0 else q / n
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def invert_pose(q, t):   # base/pose.cc:192-196
    qi = np.array([q[0], -q[1], -q[2], -q[3]])
    return qi, -(quat_R(qi) @ t)


def pair_id(a, b):
    return KMAX * b + a if a > b else KMAX * a + b


def _batch(rng, n_pairs=40, n_images=12, min_inl=15):
    ids = rng.choice(np.arange(1, 5000), n_images, replace=False).astype(np.uint32)
    pairs, seen = [], set()
    while len(pairs) < n_pairs:
        a, b = rng.integers(0, n_images, 2)
        if a != b and (min(a, b), max(a, b)) not in seen:
            seen.add((min(a, b), max(a, b)))
            pairs.append((a, b))
    pairs = np.array(pairs, np.uint32)
    m = rng.integers(0, 60, n_pairs)
    m[:4] = [0, 14, 15, 16]
    off = np.concatenate([[0], np.cumsum(m)]).astype(np.int64)
    matches = rng.integers(0, 4000, (off[-1], 2)).astype(np.uint32)
    inl = np.zeros_like(matches)
    res = np.zeros(n_pairs, RESULT_DTYPE)
    poses = np.zeros(n_pairs, POSE_DTYPE)
    for p in range(n_pairs):
        k = int(rng.integers(0, m[p] + 1)) if m[p] >= min_inl else 0
        if p % 5 == 0:
            k = min(k, 10)                                  # inliers below the gate
        res["n_inliers"][p] = k
        res["config"][p] = rng.choice([2, 3, 6, 7]) if k else 1
        sel = np.sort(rng.choice(m[p], k, replace=False)) if k else np.zeros(0, int)
        inl[off[p]:off[p] + k] = matches[off[p]:off[p + 1]][sel]
        if k and res["config"][p] != 3:
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            poses["qvec"][p], poses["tvec"][p], poses["tri_angle"][p] = q, rng.normal(size=3), 0.1
        poses["config"][p] = {6: 4}.get(int(res["config"][p]), int(res["config"][p]))
    return dict(ids=ids, pairs=pairs, off=off, matches=matches, inl=inl, res=res, poses=poses, min_inl=min_inl)


def _run(exe, tmp, b, has_pose=True):
    fin, fout = tmp / "in.bin", tmp / "out.bin"
    with open(fin, "wb") as f:
        f.write(struct.pack("<4q", len(b["pairs"]), len(b["ids"]), b["min_inl"], 1 if has_pose else 0))
        for a in (b["pairs"], b["ids"], b["off"], b["matches"], b["inl"], b["res"], b["poses"]):
            f.write(np.ascontiguousarray(a).tobytes())
    r = subprocess.run([str(exe), str(fin), str(fout)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "rows ok" in r.stdout, r.stdout + r.stderr
    raw, pos, rows = fout.read_bytes(), 0, []

    def i64():
        nonlocal pos
        v = struct.unpack_from("<q", raw, pos)[0]
        pos += 8
        return v

    def blob():
        nonlocal pos
        n = i64()
        v = raw[pos:pos + n]
        pos += n
        return v
    for _ in range(len(b["pairs"])):
        m = (i64(), i64(), i64(), blob())
        g = (i64(), i64(), i64(), blob(), i64(), blob(), blob())
        rows.append((m, g))
    assert pos == len(raw)
    return rows


def test_rows_round_trip_through_a_database_with_the_reference_schema(tmp_path):
    rng = np.random.default_rng(3)
    b = _batch(rng)
    rows = _run(_exe(tmp_path), tmp_path, b)
    db = sqlite3.connect(":memory:")
    db.execute("CREATE TABLE IF NOT EXISTS matches (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, cols INTEGER NOT NULL, data BLOB);")
    db.execute("CREATE TABLE IF NOT EXISTS two_view_geometries (pair_id INTEGER PRIMARY KEY NOT NULL, rows INTEGER NOT NULL, "
               "cols INTEGER NOT NULL, data BLOB, config INTEGER NOT NULL, F BLOB, E BLOB, H BLOB);")
    with db:   # ONE transaction for the whole batch
        db.executemany("INSERT INTO matches(pair_id, rows, cols, data) VALUES(?, ?, ?, ?);", [m for m, _ in rows])
        db.executemany("INSERT INTO two_view_geometries(pair_id, rows, cols, data, config, F, E, H) VALUES(?, ?, ?, ?, ?, ?, ?, NULL);",
                       [g for _, g in rows])
    for p, (a, c) in enumerate(b["pairs"]):
        id1, id2 = int(b["ids"][a]), int(b["ids"][c])
        swap = id1 > id2
        lo, hi = b["off"][p], b["off"][p + 1]
        # Database::ReadMatches(image_id1, image_id2)
        r, cc, data = db.execute("SELECT rows, cols, data FROM matches WHERE pair_id = ?;", (pair_id(id1, id2),)).fetchone()
        got = np.frombuffer(data or b"", np.uint32).reshape(r, cc)
        if swap:
            got = got[:, ::-1]
        exp = b["matches"][lo:hi] if hi - lo >= b["min_inl"] else np.zeros((0, 2), np.uint32)
        assert cc == 2 and (got == exp).all()
        # Database::ReadTwoViewGeometry(image_id1, image_id2)
        r, cc, data, config, F, E, H = db.execute("SELECT rows, cols, data, config, F, E, H FROM two_view_geometries WHERE pair_id = ?;",
                                                  (pair_id(id1, id2),)).fetchone()
        inl = np.frombuffer(data or b"", np.uint32).reshape(r, cc)
        q = np.frombuffer(F, np.float64) if F else np.zeros(4)
        t = np.frombuffer(E, np.float64) if E else np.zeros(3)
        if swap:                      # TwoViewGeometry::Invert on the way out
            inl = inl[:, ::-1]
            q, t = invert_pose(q, t)
        k = int(b["res"]["n_inliers"][p])
        assert H is None
        if k < b["min_inl"]:          # the gate of matching.cc:828-831: TwoViewGeometry()
            assert r == 0 and config == 0 and not F and not E
            continue
        assert config == b["poses"]["config"][p] and (inl == b["inl"][lo:lo + k]).all()
        assert np.allclose(q, b["poses"]["qvec"][p], atol=1e-15) and np.allclose(t, b["poses"]["tvec"][p], atol=1e-14)
        # and the stored bytes are what the reference's write path stores for the canonical (smaller id first) order
        eq, et = (invert_pose(b["poses"]["qvec"][p], b["poses"]["tvec"][p]) if swap else (b["poses"]["qvec"][p], b["poses"]["tvec"][p]))
        assert np.allclose(np.frombuffer(F, np.float64), eq, atol=1e-15) and np.allclose(np.frombuffer(E, np.float64), et, atol=1e-14)
        assert len(F) == 32 and len(E) == 24
    assert db.execute("SELECT COUNT(*) FROM two_view_geometries WHERE rows > 0;").fetchone()[0] == int((b["res"]["n_inliers"] >= 15).sum())


def test_rows_without_relative_pose_keep_the_constructor_values(tmp_path):
    rng = np.random.default_rng(5)
    b = _batch(rng, n_pairs=12)
    rows = _run(_exe(tmp_path), tmp_path, b, has_pose=False)
    for p, (_, g) in enumerate(rows):
        if b["res"]["n_inliers"][p] >= 15:
            assert g[4] == b["res"]["config"][p] and np.frombuffer(g[6], np.float64).tolist() == [0, 0, 0]
            q = np.frombuffer(g[5], np.float64)
            assert (np.abs(q) == 0).all()
