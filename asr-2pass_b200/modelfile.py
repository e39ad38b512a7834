"""Model-directory I/O for the B200 Paraformer path.

The directory layout is the reference's (onnxruntime/include/com-define.h:52-88,
onnxruntime/src/offline-stream.cpp:60-87): `am.mvn`, `config.yaml`, `tokens.json` (+ the reference's
`model.onnx`, which this implementation does not read) plus ONE extra file, `model.b200pf`: a flat
little-endian tensor file holding the fp32 parameters under their upstream FunASR state_dict names
(SURVEY.md Appendix B).  The C++ loader (csrc/model_file.cpp) reads exactly this format.

    char  magic[8]  = "B2PFWTS1"
    u32   n_cfg ;  n_cfg  x { char key[32]; f64 value }
    u32   n_tens;  n_tens x { char name[96]; u32 ndim; u64 dims[4]; u64 offset; u64 nbytes }
    ... zero padding to a 256-byte boundary ..., then tensor data (fp32), each 256-byte aligned;
    `offset` is absolute from the start of the file.
"""
import json
import os
import struct

import numpy as np

MAGIC = b"B2PFWTS1"
WEIGHT_FILE = "model.b200pf"


def _align(x, a=256):
    return (x + a - 1) // a * a


def write_weights(path, cfg: dict, tensors: dict):
    names = list(tensors.keys())
    hdr_size = 8 + 4 + len(cfg) * 40 + 4 + len(names) * (96 + 4 + 32 + 8 + 8)
    off = _align(hdr_size)
    recs = []
    for n in names:
        a = np.ascontiguousarray(tensors[n], dtype=np.float32)
        assert a.ndim <= 4 and len(n.encode()) < 96, n
        recs.append((n, a, off))
        off = _align(off + a.nbytes)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(cfg)))
        for k, v in cfg.items():
            f.write(struct.pack("<32sd", k.encode(), float(v)))
        f.write(struct.pack("<I", len(recs)))
        for n, a, o in recs:
            dims = list(a.shape) + [0] * (4 - a.ndim)
            f.write(struct.pack("<96sI4QQQ", n.encode(), a.ndim, *dims, o, a.nbytes))
        for n, a, o in recs:
            f.seek(o)
            f.write(a.tobytes())
        f.truncate(off)


def read_weights(path):
    """-> (cfg dict, {name: np.ndarray fp32}) ; used by tests to feed the oracle the same bytes."""
    with open(path, "rb") as f:
        assert f.read(8) == MAGIC
        (n_cfg,) = struct.unpack("<I", f.read(4))
        cfg = {}
        for _ in range(n_cfg):
            k, v = struct.unpack("<32sd", f.read(40))
            cfg[k.rstrip(b"\0").decode()] = v
        (n_t,) = struct.unpack("<I", f.read(4))
        recs = []
        for _ in range(n_t):
            r = struct.unpack("<96sI4QQQ", f.read(96 + 4 + 32 + 16))
            recs.append(r)
        out = {}
        for r in recs:
            name = r[0].rstrip(b"\0").decode()
            ndim, dims, off, nbytes = r[1], r[2:6], r[6], r[7]
            f.seek(off)
            out[name] = np.frombuffer(f.read(nbytes), dtype=np.float32).reshape(dims[:ndim]).copy()
    return cfg, out


def write_am_mvn(path, means, vars_):
    """Kaldi-nnet text as parsed by Paraformer::LoadCmvn (paraformer.cpp:325-360): the line after
    `<AddShift>` / `<Rescale>` starts with `<LearnRateCoef>` and tokens [3 .. size-2] are the values."""
    n = len(means)

    def row(v):
        return "<LearnRateCoef> 0 [ " + " ".join(repr(float(np.float32(x))) for x in v) + " ]"

    with open(path, "w") as f:
        f.write("<Nnet>\n")
        f.write(f"<Splice> {n} {n}\n[ 0 ]\n")
        f.write(f"<AddShift> {n} {n}\n{row(means)}\n")
        f.write(f"<Rescale> {n} {n}\n{row(vars_)}\n")
        f.write("</Nnet>\n")


def read_am_mvn(path):
    means, vars_ = [], []
    with open(path) as f:
        lines = f.read().split("\n")
    i = 0
    while i < len(lines):
        it = lines[i].split()
        if it and it[0] in ("<AddShift>", "<Rescale>") and i + 1 < len(lines):
            nxt = lines[i + 1].split()
            if nxt and nxt[0] == "<LearnRateCoef>":
                vals = [float(x) for x in nxt[3:-1]]
                (means if it[0] == "<AddShift>" else vars_).extend(vals)
                i += 1
        i += 1
    return np.asarray(means, np.float32), np.asarray(vars_, np.float32)


def write_model_dir(path, cfg: dict, tensors: dict, means, vars_, tokens, lang="zh-cn", fs=16000):
    os.makedirs(path, exist_ok=True)
    write_weights(os.path.join(path, WEIGHT_FILE), cfg, tensors)
    write_am_mvn(os.path.join(path, "am.mvn"), means, vars_)
    with open(os.path.join(path, "tokens.json"), "w", encoding="utf-8") as f:
        json.dump(list(tokens), f, ensure_ascii=False)
    with open(os.path.join(path, "config.yaml"), "w") as f:
        f.write("frontend: wav_frontend\nfrontend_conf:\n  fs: %d\n  window: hamming\n  n_mels: 80\n"
                "  frame_length: 25\n  frame_shift: 10\n  lfr_m: 7\n  lfr_n: 6\n" % fs)
        if lang:
            f.write("lang: %s\n" % lang)
