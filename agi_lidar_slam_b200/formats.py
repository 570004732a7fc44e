"""Wire / disk formats at the two ends of the path.

* PCD (binary) — the map hand-over of the relocalisation mode: src/laserMapping.cpp:805-835 saves
  `ikdtree.flatten()` with `pcl::PCDWriter::writeBinary` as PCD/GlobalMap_ikdtree.pcd, src/laserMapping_re.cpp:341-352
  loads it with `pcl::io::loadPCDFile<PointType>` and hands it to `ikdtree.Build`.  PointType = pcl::PointXYZINormal.
  PCL is not in this image: the layout below is PCL's published PCD v0.7 format ([ext], unpinned by the reference).
* odometry — nav_msgs::Odometry pose + 6x6 covariance as publish_odometry fills it (src/laserMapping.cpp:537-553).
"""
from __future__ import annotations

import numpy as np

# registration order of pcl::PointXYZINormal (what PCDWriter::writeBinary packs, 32 bytes per point) and the float
# index of each field inside the 48-byte in-memory record {x,y,z,_}{nx,ny,nz,_}{intensity,curvature,_,_}
PCD_FIELDS = ("x", "y", "z", "intensity", "normal_x", "normal_y", "normal_z", "curvature")
_RECORD_INDEX = {"x": 0, "y": 1, "z": 2, "normal_x": 4, "normal_y": 5, "normal_z": 6, "intensity": 8, "curvature": 9}


def write_pcd_binary(path, records: np.ndarray) -> None:
    """records: (n,12) float32 PointXYZINormal records, or (n,3)/(n,4) xyz[/intensity] (other fields zero)."""
    pts = np.asarray(records, np.float32)
    n = pts.shape[0]
    packed = np.zeros((n, len(PCD_FIELDS)), np.float32)
    if pts.shape[1] == 12:
        for k, f in enumerate(PCD_FIELDS):
            packed[:, k] = pts[:, _RECORD_INDEX[f]]
    else:
        packed[:, :3] = pts[:, :3]
        if pts.shape[1] >= 4:
            packed[:, 3] = pts[:, 3]
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS " + " ".join(PCD_FIELDS) + "\nSIZE " +
           " ".join(["4"] * 8) + "\nTYPE " + " ".join(["F"] * 8) + "\nCOUNT " + " ".join(["1"] * 8) +
           f"\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")
    with open(path, "wb") as fh:
        fh.write(hdr.encode("ascii"))
        fh.write(packed.tobytes())


def read_pcd(path) -> np.ndarray:
    """Any binary / ascii PCD with float32 x y z (and optionally the other PointXYZINormal fields; `_` padding fields
    and fields of other names are skipped) -> (n,12) float32 PointXYZINormal records, ready for lio_map_build."""
    with open(path, "rb") as fh:
        raw = fh.read()
    hdr, pos = {}, 0
    while True:
        end = raw.index(b"\n", pos)
        line = raw[pos:end].decode("ascii", "replace").strip()
        pos = end + 1
        if not line or line.startswith("#"):
            continue
        key, _, val = line.partition(" ")
        hdr[key.upper()] = val.split()
        if key.upper() == "DATA":
            break
    fields = hdr["FIELDS"]
    sizes = [int(v) for v in hdr["SIZE"]]
    types = hdr["TYPE"]
    counts = [int(v) for v in hdr.get("COUNT", ["1"] * len(fields))]
    n = int(hdr["POINTS"][0]) if "POINTS" in hdr else int(hdr["WIDTH"][0]) * int(hdr.get("HEIGHT", ["1"])[0])
    mode = hdr["DATA"][0].lower()
    out = np.zeros((n, 12), np.float32)
    out[:, 3] = 1.0
    np_type = {("F", 4): "<f4", ("F", 8): "<f8", ("U", 1): "u1", ("U", 2): "<u2", ("U", 4): "<u4", ("I", 1): "i1",
               ("I", 2): "<i2", ("I", 4): "<i4"}
    if mode == "binary":
        dt = np.dtype([(f"f{k}", np_type[(types[k], sizes[k])], (counts[k],)) for k in range(len(fields))])
        rec = np.frombuffer(raw, dt, count=n, offset=pos)
        for k, f in enumerate(fields):
            if f in _RECORD_INDEX:
                out[:, _RECORD_INDEX[f]] = rec[f"f{k}"][:, 0].astype(np.float32)
    elif mode == "ascii":
        tab = np.loadtxt(raw[pos:].decode("ascii").splitlines(), dtype=np.float64, ndmin=2)
        col = 0
        for k, f in enumerate(fields):
            if f in _RECORD_INDEX:
                out[:, _RECORD_INDEX[f]] = tab[:, col].astype(np.float32)
            col += counts[k]
    else:
        raise ValueError(f"PCD DATA {mode} is not supported (binary_compressed needs LZF)")
    return out


def odometry_message(x: np.ndarray, P: np.ndarray):
    """publish_odometry (src/laserMapping.cpp:537-553): position, orientation (x,y,z,w) and the 6x6 pose covariance in
    ROS order (translation first) taken from the filter's (pos 0-2, rot 3-5) blocks with the reference's index swap
    k = i < 3 ? i + 3 : i - 3."""
    x = np.asarray(x, np.float64)
    P = np.asarray(P, np.float64).reshape(24, 24)
    cov = np.zeros(36)
    for i in range(6):
        k = i + 3 if i < 3 else i - 3
        cov[i * 6 + 0] = P[k, 3]
        cov[i * 6 + 1] = P[k, 4]
        cov[i * 6 + 2] = P[k, 5]
        cov[i * 6 + 3] = P[k, 0]
        cov[i * 6 + 4] = P[k, 1]
        cov[i * 6 + 5] = P[k, 2]
    return dict(position=x[0:3].copy(), orientation_xyzw=np.array([x[4], x[5], x[6], x[3]]), covariance=cov)
