"""A binary batch format for the per-frame detections ("a binary batch format would remove JSON from the
loop").

The reference reads one ``<token>.json`` per frame and per class model, and builds the ``(max_objects, 11)`` box array
row by row in Python (det3d/datasets/nuscenes/nuscenes.py:207-293). A detection file holds the same information for
many frames in flat arrays, so that a BATCH of frame pairs is packed with a handful of numpy slice copies:

    header   magic "SHDB0002", little-endian u32 counts (frames, classes, attributes), u64 rows
    tables   class names, attribute names (u16 length + utf-8 each)
    frames   token[frames] (u16 length + utf-8), then int32 prev_index (-1 = first frame of a scene), int64 timestamp,
             int64 prev_timestamp (microseconds, nuScenes frame info), int64 row_begin, int32 row_count
    rows     float64 [rows][14]: the ``det_path`` box in the SENSOR frame of its own sweep - translation(3) size(3)
             rotation(4: w,x,y,z) velocity(2) - then detection_score and yaw: what the head's (M, 11) boxes are packed from
             float64 [rows][12]: the ``cls_info`` entry of the same detection in the GLOBAL frame - translation(3)
             size(3) rotation(4) velocity(2): what the eval loop emits and the tracker matches on
             (tools/nusc_shasta/eval.py:126-181; the reference keeps the two in separate directories,
             configs det_path = '.../sensor_individual_frames', preprocessing/filter_track_types.py)
             uint8 class id [rows], uint8 attribute id [rows]   (255 = no attribute)
             uint32 extra_offset [rows + 1] + utf-8 blob: JSON object of any further ``cls_info`` keys of a row

Rows keep the order of the frame's detection list (all classes), so the ``keep`` indices of ``formats.pack_detections``
are reproduced; values stay float64 so that the emitted annotations are the input's, and ``yaw`` is computed ONCE at
write time with the same scalar routine as the JSON path (``formats.quaternion_yaw``), which makes the packed arrays
bit-identical to ``formats.frame_pair_example`` on the JSON input (tests/test_detfile.py).
"""
import json as _json
import random as _random
import struct

import numpy as np

from . import formats

MAGIC = b"SHDB0002"
ROW = 14
CLS_ROW = 12
_CLS_KEYS = ("sample_token", "translation", "size", "rotation", "velocity", "detection_name", "detection_score",
             "attribute_name")
NO_ATTR = 255


def _put_str(out, s):
    b = s.encode("utf-8")
    out.append(struct.pack("<H", len(b)))
    out.append(b)


def write_detection_file(path, frames):
    """``frames``: list of ``{token, prev_token ('' at a scene start), timestamp, prev_timestamp, dets, cls}`` in the
    per-frame detection JSON layout (see ``pipeline.run_class_sequence``). ``prev_token`` must name a frame of the same
    file (or be '')."""
    index = {f["token"]: i for i, f in enumerate(frames)}
    classes, attrs = [], []
    rows, cls_rows, cls_id, attr_id, extras = [], [], [], [], []
    prev_index, ts, pts, begin, count = [], [], [], [], []
    for f in frames:
        prev_index.append(index[f["prev_token"]] if f["prev_token"] != "" else -1)
        ts.append(int(f["timestamp"]))
        pts.append(int(f["prev_timestamp"]))
        begin.append(len(rows))
        count.append(len(f["dets"]))
        for b, info in zip(f["dets"], f["cls"]):
            name = info["detection_name"]
            if name not in classes:
                classes.append(name)
            attr = info.get("attribute_name")
            if attr is not None and attr not in attrs:
                attrs.append(attr)
            rows.append(list(b[:12]) + [info["detection_score"], formats.quaternion_yaw(np.array(b[6:10]))[0]])
            # cls_info fields the entry does not carry default to the det_path box (synthetic scenes)
            cls_rows.append(list(info.get("translation", b[0:3])[:3]) + list(info.get("size", b[3:6])[:3]) +
                            list(info.get("rotation", b[6:10])[:4]) + list(info.get("velocity", b[10:12])[:2]))
            cls_id.append(classes.index(name))
            attr_id.append(NO_ATTR if attr is None else attrs.index(attr))
            extra = {k: v for k, v in info.items() if k not in _CLS_KEYS}
            extras.append(_json.dumps(extra).encode("utf-8") if extra else b"")
    if len(classes) > 255 or len(attrs) > 254:
        raise ValueError("too many class / attribute names for the one-byte ids")
    out = [MAGIC, struct.pack("<IIIQ", len(frames), len(classes), len(attrs), len(rows))]
    for s in classes + attrs + [f["token"] for f in frames]:
        _put_str(out, s)
    head = b"".join(out)
    head += b"\0" * (-len(head) % 8)          # the arrays start 8-byte aligned
    with open(path, "wb") as fh:
        fh.write(head)
        fh.write(np.asarray(prev_index, "<i4").tobytes())
        fh.write(b"\0" * (-4 * len(frames) % 8))
        fh.write(np.asarray(ts, "<i8").tobytes())
        fh.write(np.asarray(pts, "<i8").tobytes())
        fh.write(np.asarray(begin, "<i8").tobytes())
        fh.write(np.asarray(count, "<i4").tobytes())
        fh.write(b"\0" * (-4 * len(frames) % 8))
        fh.write(np.asarray(rows, "<f8").reshape(-1, ROW).tobytes())
        fh.write(np.asarray(cls_rows, "<f8").reshape(-1, CLS_ROW).tobytes())
        fh.write(np.asarray(cls_id, "u1").tobytes())
        fh.write(np.asarray(attr_id, "u1").tobytes())
        fh.write(b"\0" * (-2 * len(rows) % 8))
        offs = np.zeros(len(rows) + 1, "<u4")
        if extras:
            offs[1:] = np.cumsum([len(e) for e in extras])
        fh.write(offs.tobytes())
        fh.write(b"".join(extras))


class DetectionFile:
    """Read side: arrays are views of one memory map."""

    def __init__(self, path):
        buf = np.memmap(path, dtype="u1", mode="r")
        if bytes(buf[:8]) != MAGIC:
            raise ValueError("%s is not a detection file (bad magic)" % path)
        nf, nc, na, nr = struct.unpack("<IIIQ", bytes(buf[8:28]))
        pos = 28

        def get_str():
            nonlocal pos
            (n,) = struct.unpack("<H", bytes(buf[pos:pos + 2]))
            s = bytes(buf[pos + 2:pos + 2 + n]).decode("utf-8")
            pos += 2 + n
            return s

        self.classes = [get_str() for _ in range(nc)]
        self.attributes = [get_str() for _ in range(na)]
        self.tokens = [get_str() for _ in range(nf)]
        pos += -pos % 8

        def take(dtype, n, pad8=False):
            nonlocal pos
            nbytes = np.dtype(dtype).itemsize * n
            if pos + nbytes > buf.size:
                raise ValueError("%s is truncated" % path)
            a = buf[pos:pos + nbytes].view(dtype)
            pos += nbytes
            if pad8:
                pos += -nbytes % 8
            return a

        self.prev_index = take("<i4", nf, pad8=True)
        self.timestamp = take("<i8", nf)
        self.prev_timestamp = take("<i8", nf)
        self.row_begin = take("<i8", nf)
        self.row_count = take("<i4", nf, pad8=True)
        self.rows = take("<f8", nr * ROW).reshape(nr, ROW)
        self.cls_rows = take("<f8", nr * CLS_ROW).reshape(nr, CLS_ROW)
        self.cls_id = take("u1", nr)
        self.attr_id = take("u1", nr)
        pos += -2 * nr % 8
        self.extra_offset = take("<u4", nr + 1)
        self.extra_blob = take("u1", int(self.extra_offset[nr]) if nr else 0)
        self.n_frames = nf
        self._index = {t: i for i, t in enumerate(self.tokens)}

    def frame_index(self, token):
        return self._index[token]

    def time_diff(self, i):
        """eval-loop time difference of frame pair i, computed like the reference's dataset (nuscenes.py:209)."""
        return 1e-6 * int(self.timestamp[i]) - 1e-6 * int(self.prev_timestamp[i])

    def _select(self, i, det_type):
        """Row numbers (absolute) and in-frame indices of frame i's detections of the requested classes."""
        b, n = int(self.row_begin[i]), int(self.row_count[i])
        local = np.arange(n)
        if det_type is not None:
            ids = [self.classes.index(c) for c in det_type if c in self.classes]
            local = local[np.isin(self.cls_id[b:b + n], ids)]
        return b + local, local

    def pack(self, i, max_objects, time_diff, det_type=None, rng=None):
        """``formats.pack_detections`` for frame i: ``(boxes (max_objects, 11) float64, keep, absolute row numbers of
        the kept detections, real count)``."""
        rng = rng if rng is not None else _random
        boxes = np.zeros((max_objects, 11))
        if int(self.row_count[i]) == 0:
            return boxes, list(range(max_objects)), np.zeros(0, np.int64), 0
        absolute, local = self._select(i, det_type)
        if len(local) > max_objects:
            idx = rng.sample(range(len(local)), max_objects)
            idx.sort()
            absolute, local = absolute[idx], local[idx]
        n = len(local)
        if n > 0:
            r = self.rows[absolute]
            boxes[:n, 0:6] = r[:, 0:6]
            boxes[:n, 6] = r[:, 13]
            boxes[:n, 7:9] = r[:, 10:12]
            boxes[:n, 9] = time_diff
            boxes[:n, 10] = r[:, 12]
        return boxes, [int(k) for k in local], absolute, n

    def scenes(self):
        """[(first frame index, number of frames)] of the scenes: a frame without a previous frame starts one, and a
        scene's frames must follow each other in time order (what ``multiclass.run_sequence_batch`` shards by)."""
        out = []
        for i in range(self.n_frames):
            p = int(self.prev_index[i])
            if p < 0:
                out.append([i, 1])
            elif out and p == i - 1 and out[-1][0] <= p:
                out[-1][1] += 1
            else:
                raise ValueError("frame %d (%s) does not follow its previous frame" % (i, self.tokens[i]))
        return [tuple(x) for x in out]

    def frame_pair_batch(self, frame_indices, max_objects, det_type=None, rng=None, rng_for=None):
        """The ``example`` arrays of a batch of frame pairs: float32 ``det_boxes`` / ``prev_det_boxes`` (B, M, 11),
        int32 ``n_det`` / ``n_prev``, and per pair the ``keep`` lists and absolute row numbers (for ``cls_info``).
        The previous frame of pair i is ``prev_index[i]`` (first frame of a scene: zeros, like the reference).
        ``rng_for(i)``: optional per-frame-pair generator for the sub-sampling of over-full frames, so that the choice
        does not depend on how the frame pairs are batched or sharded."""
        B, M = len(frame_indices), max_objects
        det = np.zeros((B, M, 11), np.float32)
        prev = np.zeros((B, M, 11), np.float32)
        n_det, n_prev = np.zeros(B, np.int32), np.zeros(B, np.int32)
        keep, prev_keep, rows, prev_rows = [], [], [], []
        for j, i in enumerate(frame_indices):
            dt = self.time_diff(i)
            if rng_for is not None:
                rng = rng_for(i)
            p = int(self.prev_index[i])
            if p < 0:
                pk, pr = list(range(M)), np.zeros(0, np.int64)
            else:
                pb, pk, pr, n_prev[j] = self.pack(p, M, dt, det_type, rng)
                prev[j] = pb
            cb, ck, cr, n_det[j] = self.pack(i, M, dt, det_type, rng)
            det[j] = cb
            keep.append(ck), prev_keep.append(pk), rows.append(cr), prev_rows.append(pr)
        return {"det_boxes": det, "prev_det_boxes": prev, "n_det": n_det, "n_prev": n_prev, "keep": keep,
                "prev_keep": prev_keep, "rows": rows, "prev_rows": prev_rows}

    def cls_info(self, absolute_rows, token):
        """The detection dicts of the given rows, in the layout the eval loop emits (``cp_<split>.json`` entries):
        the GLOBAL-frame ``cls_info`` fields, not the sensor-frame box the head's input is packed from."""
        out = []
        for r in absolute_rows:
            r = int(r)
            v = self.cls_rows[r]
            d = {"sample_token": token, "translation": v[0:3].tolist(), "size": v[3:6].tolist(),
                 "rotation": v[6:10].tolist(), "velocity": v[10:12].tolist(),
                 "detection_name": self.classes[int(self.cls_id[r])], "detection_score": float(self.rows[r, 12])}
            if int(self.attr_id[r]) != NO_ATTR:
                d["attribute_name"] = self.attributes[int(self.attr_id[r])]
            lo, hi = int(self.extra_offset[r]), int(self.extra_offset[r + 1])
            if hi > lo:
                d.update(_json.loads(bytes(self.extra_blob[lo:hi]).decode("utf-8")))
            out.append(d)
        return out

    def frames(self):
        """Back to the JSON-layout frame list ``write_detection_file`` takes (round trip)."""
        res = []
        for i, token in enumerate(self.tokens):
            b, n = int(self.row_begin[i]), int(self.row_count[i])
            p = int(self.prev_index[i])
            res.append({"token": token, "prev_token": "" if p < 0 else self.tokens[p],
                        "timestamp": int(self.timestamp[i]), "prev_timestamp": int(self.prev_timestamp[i]),
                        "dets": [self.rows[r, 0:12].tolist() for r in range(b, b + n)],
                        "cls": self.cls_info(range(b, b + n), token)})
        return res


def frames_from_json_dirs(det_path, cls_info_path, frame_info, tokens):
    """The reference's on-disk layout (nuscenes.py:198-217, 255-262): ``<det_path>/<token>.json`` = list of 12-float
    boxes, ``<cls_info_path>/<token>.json`` = list of detection dicts, ``frame_info[token] = {prev, timestamp,
    prev_timestamp}``. ``tokens``: the frames to convert, scene by scene in time order; a ``prev`` outside ``tokens``
    counts as a scene start (the reference's ``get_frame_idx(prev_token) is None`` case). Returns the frame list
    ``write_detection_file`` takes."""
    import json
    import os
    known = set(tokens)
    frames = []
    for token in tokens:
        with open(os.path.join(det_path, token + ".json")) as fh:
            dets = json.load(fh)
        with open(os.path.join(cls_info_path, token + ".json")) as fh:
            cls = json.load(fh)
        info = frame_info[token]
        prev = info["prev"] if info["prev"] in known else ""
        frames.append({"token": token, "prev_token": prev, "timestamp": info["timestamp"],
                       "prev_timestamp": info["prev_timestamp"], "dets": dets, "cls": cls})
    return frames
