"""Torch7 binary serialisation (`torch.save` / `torch.load`, torch7 File.lua) -- reader, a minimal writer for tests, and the
mapping of a DenseCap checkpoint (train.lua:157-185: {opt, iter, ..., model = DenseCapModel}) onto the parameter names of
`dc_load_weights` (replaces `torch.load(opt.checkpoint).model`, run_model.lua:146-149).

Format (little-endian, "binary" mode, 64-bit longs):
  object := int32 type ; payload
    0 nil | 1 number: float64 | 2 string: int32 len, bytes | 5 boolean: int32
    3 table : int32 index ; [first time] int32 n ; n x (key object, value object)
    4 torch : int32 index ; [first time] string "V 1" ; string class name ; class payload
  torch.XTensor  payload: int32 ndim ; int64 size[ndim] ; int64 stride[ndim] ; int64 storage offset (1-based) ; storage object
  torch.XStorage payload: int64 n ; n raw elements
  any other class (nn.*, DenseCapModel, ...): one object = the table of its fields
`index` implements reference sharing: an index seen before stands for the object already read.

Pure host logic (numpy only); nothing here runs on the hot path."""
from __future__ import annotations

import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
TYPE_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION, TYPE_RECUR_FUNCTION = 6, 7, 8

_DTYPES = {"Float": np.float32, "Double": np.float64, "Long": np.int64, "Int": np.int32, "Short": np.int16, "Byte": np.uint8,
           "Char": np.int8, "Half": np.float16, "Cuda": np.float32, "CudaLong": np.int64, "CudaDouble": np.float64, "CudaHalf": np.float16}


class T7Object:
    """A torch class instance that has no special payload: class name + the table of its fields."""

    def __init__(self, classname, fields=None):
        self.classname = classname
        self.fields = fields if fields is not None else {}

    def __getitem__(self, k):
        return self.fields[k]

    def get(self, k, default=None):
        return self.fields.get(k, default) if isinstance(self.fields, dict) else default

    def __repr__(self):
        return "T7Object(%s, %d fields)" % (self.classname, len(self.fields) if hasattr(self.fields, "__len__") else 0)


def _tensor_kind(classname):
    """'torch.FloatTensor' -> ('Float', 'Tensor'); None for other classes."""
    if not classname.startswith("torch."):
        return None
    name = classname[len("torch."):]
    for suffix in ("Tensor", "Storage"):
        if name.endswith(suffix) and name[:-len(suffix)] in _DTYPES:
            return name[:-len(suffix)], suffix
    return None


class Reader:
    def __init__(self, data: bytes):
        self.b = memoryview(data)
        self.p = 0
        self.objects = {}

    def _need(self, n):
        if n < 0 or self.p + n > len(self.b):
            raise ValueError("t7: truncated or corrupt stream (need %d bytes at offset %d of %d)" % (n, self.p, len(self.b)))

    def _unpack(self, fmt, n):
        self._need(n)
        v = struct.unpack_from("<" + fmt, self.b, self.p)
        self.p += n
        return v[0]

    def int32(self):
        return self._unpack("i", 4)

    def int64(self):
        return self._unpack("q", 8)

    def string(self):
        n = self.int32()
        self._need(n)
        s = bytes(self.b[self.p:self.p + n])
        self.p += n
        return s.decode("utf-8", "replace")

    def obj(self):
        t = self.int32()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = self._unpack("d", 8)
            return int(v) if float(v).is_integer() and abs(v) < 2 ** 53 else v
        if t == TYPE_BOOLEAN:
            return self.int32() != 0
        if t == TYPE_STRING:
            return self.string()
        if t == TYPE_TABLE:
            idx = self.int32()
            if idx in self.objects:
                return self.objects[idx]
            out = {}
            self.objects[idx] = out
            n = self.int32()
            for _ in range(n):
                k = self.obj()
                out[k] = self.obj()
            return out
        if t == TYPE_TORCH:
            idx = self.int32()
            if idx in self.objects:
                return self.objects[idx]
            version = self.string()
            classname = self.string() if version.startswith("V ") else version      # pre-versioning files store the class name first
            kind = _tensor_kind(classname)
            if kind and kind[1] == "Storage":
                n = self.int64()
                dt = np.dtype(_DTYPES[kind[0]])
                self._need(n * dt.itemsize)
                arr = np.frombuffer(self.b, dtype=dt, count=n, offset=self.p).copy()
                self.p += n * dt.itemsize
                self.objects[idx] = arr
                return arr
            if kind and kind[1] == "Tensor":
                nd = self.int32()
                size = [self.int64() for _ in range(nd)]
                stride = [self.int64() for _ in range(nd)]
                off = self.int64() - 1
                storage = self.obj()
                dt = np.dtype(_DTYPES[kind[0]])
                if storage is None or nd == 0:
                    arr = np.zeros((0,), dt)
                else:
                    last = off + sum((sz - 1) * st for sz, st in zip(size, stride)) if all(sz > 0 for sz in size) else off
                    if off < 0 or min(size) < 0 or min(stride) < 0 or (all(sz > 0 for sz in size) and last >= storage.size):
                        raise ValueError("t7: tensor view (size %s, stride %s, offset %d) exceeds its storage of %d elements"
                                         % (size, stride, off + 1, storage.size))
                    arr = np.lib.stride_tricks.as_strided(storage[off:], shape=size, strides=[s * dt.itemsize for s in stride]).copy()
                self.objects[idx] = arr
                return arr
            o = T7Object(classname)
            self.objects[idx] = o
            o.fields = self.obj()
            return o
        if t in (TYPE_FUNCTION, LEGACY_TYPE_RECUR_FUNCTION, TYPE_RECUR_FUNCTION):
            raise ValueError("t7: serialised Lua functions are not supported")
        raise ValueError("t7: unknown type tag %d at byte %d" % (t, self.p - 4))


def load(path):
    with open(path, "rb") as f:
        return Reader(f.read()).obj()


def loads(data: bytes):
    return Reader(data).obj()


# ---------------------------------------------------------------------------------------------------------
# minimal writer (tests, and exporting synthetic weights in the reference's container)
# ---------------------------------------------------------------------------------------------------------
class Writer:
    def __init__(self):
        self.out = bytearray()
        self.next_index = 1
        self.seen = {}

    def int32(self, v):
        self.out += struct.pack("<i", int(v))

    def int64(self, v):
        self.out += struct.pack("<q", int(v))

    def string(self, s):
        b = s.encode("utf-8")
        self.int32(len(b))
        self.out += b

    def _index(self, o):
        """Returns True when the object was written before (only its index is emitted)."""
        key = id(o)
        if key in self.seen:
            self.int32(self.seen[key][0])
            return True
        self.seen[key] = (self.next_index, o)          # keep `o` alive so ids stay unique
        self.int32(self.next_index)
        self.next_index += 1
        return False

    def obj(self, o):
        if o is None:
            self.int32(TYPE_NIL)
        elif isinstance(o, bool):
            self.int32(TYPE_BOOLEAN); self.int32(1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.int32(TYPE_NUMBER); self.out += struct.pack("<d", float(o))
        elif isinstance(o, str):
            self.int32(TYPE_STRING); self.string(o)
        elif isinstance(o, np.ndarray):
            names = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.int64): "Long", np.dtype(np.int32): "Int",
                     np.dtype(np.uint8): "Byte"}
            kind = names[o.dtype]
            self.int32(TYPE_TORCH)
            if self._index(o):
                return
            self.string("V 1"); self.string("torch.%sTensor" % kind)
            a = np.ascontiguousarray(o)
            self.int32(a.ndim)
            for s in a.shape:
                self.int64(s)
            for s in a.strides:
                self.int64(s // a.itemsize)
            self.int64(1)
            self.int32(TYPE_TORCH)
            self.int32(self.next_index); self.next_index += 1
            self.string("V 1"); self.string("torch.%sStorage" % kind)
            self.int64(a.size)
            self.out += a.tobytes()
        elif isinstance(o, T7Object):
            self.int32(TYPE_TORCH)
            if self._index(o):
                return
            self.string("V 1"); self.string(o.classname)
            self.obj(o.fields)
        elif isinstance(o, (list, tuple)):
            self.obj({i + 1: v for i, v in enumerate(o)})
        elif isinstance(o, dict):
            self.int32(TYPE_TABLE)
            if self._index(o):
                return
            self.int32(len(o))
            for k, v in o.items():
                self.obj(k); self.obj(v)
        else:
            raise TypeError("t7 writer: unsupported %r" % type(o))


def dumps(o) -> bytes:
    w = Writer()
    w.obj(o)
    return bytes(w.out)


def save(o, path):
    with open(path, "wb") as f:
        f.write(dumps(o))


# ---------------------------------------------------------------------------------------------------------
# DenseCap checkpoint -> dc_load_weights names
# ---------------------------------------------------------------------------------------------------------
def _array_items(t):
    """Values of a Lua array-like table in index order (1..n)."""
    if isinstance(t, dict):
        keys = sorted(k for k in t if isinstance(k, int))
        return [t[k] for k in keys]
    return list(t or [])


def _walk_modules(m, out):
    """Depth-first, in `modules` order (nn.Sequential / ConcatTable / ParallelTable containers)."""
    if isinstance(m, T7Object):
        f = m.fields if isinstance(m.fields, dict) else {}
        if isinstance(f.get("weight"), np.ndarray) and f["weight"].size:
            out.append(m)
        for child in _array_items(f.get("modules")):
            _walk_modules(child, out)
    return out


def _field(o, *path):
    for k in path:
        if isinstance(o, T7Object):
            o = o.fields
        if not isinstance(o, dict) or k not in o:
            raise KeyError("checkpoint: missing field %r" % (k,))
        o = o[k]
    return o


VGG16_CONV_NAMES = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
                    "conv5_1", "conv5_2", "conv5_3"]


def extract_densecap_weights(checkpoint):
    """checkpoint: what `load()` returns for a file written by train.lua:184 (or the DenseCapModel object itself).
    Returns (weights: name -> float32 array in the layouts `dc_load_weights` takes, info: dict with vocab_size, idx_to_token, opt).
    Structure followed: DenseCapModel.lua:40-125 (nets.conv_net1/conv_net2/localization_layer/recog_base/objectness_branch/
    box_reg_branch/language_model), LocalizationLayer.lua:609-690 (rpn: conv, box conv, score conv in `modules` order),
    LanguageModel.lua:26-71 (image_encoder, lookup_table, rnn = LSTM ... Linear)."""
    model = checkpoint["model"] if isinstance(checkpoint, dict) and "model" in checkpoint else checkpoint
    nets = _field(model, "nets")
    w = {}
    convs = _walk_modules(nets["conv_net1"], []) + _walk_modules(nets["conv_net2"], [])
    convs = [m for m in convs if m.fields["weight"].ndim in (2, 4)]
    if len(convs) != 13:
        raise ValueError("checkpoint: expected the 13 VGG-16 convolutions in conv_net1 + conv_net2, found %d" % len(convs))
    cin = 3
    for name, m in zip(VGG16_CONV_NAMES, convs):
        wt = np.asarray(m.fields["weight"], np.float32)
        cout = wt.shape[0]
        w[name + ".weight"] = wt.reshape(cout, cin, 3, 3)        # nn.SpatialConvolutionMM keeps (Cout, Cin*3*3)
        w[name + ".bias"] = np.asarray(m.fields["bias"], np.float32).reshape(cout)
        cin = cout
    rpn = _walk_modules(_field(nets["localization_layer"], "nets", "rpn"), [])
    if len(rpn) != 3:
        raise ValueError("checkpoint: expected 3 convolutions in the RPN, found %d" % len(rpn))
    # `modules` order (LocalizationLayer.lua:609-690): 3x3 conv, then box_branch's 1x1 conv, then rpn_branch's 1x1 score conv
    F = rpn[0].fields["weight"].shape[0]
    for key, m, kk in (("rpn.conv", rpn[0], 3), ("rpn.box", rpn[1], 1), ("rpn.score", rpn[2], 1)):
        wt = np.asarray(m.fields["weight"], np.float32)
        rows = wt.shape[0]
        in_ch = cin if kk == 3 else F
        if wt.size != rows * in_ch * kk * kk:
            raise ValueError("checkpoint: %s weight has %d elements, expected %d x %d x %d x %d" % (key, wt.size, rows, in_ch, kk, kk))
        w[key + ".weight"] = wt.reshape(rows, in_ch, kk, kk)
        w[key + ".bias"] = np.asarray(m.fields["bias"], np.float32).reshape(rows)
    if w["rpn.box.weight"].shape[0] != 2 * w["rpn.score.weight"].shape[0]:
        raise ValueError("checkpoint: RPN box / score heads have inconsistent anchor counts")
    fcs = _walk_modules(nets["recog_base"], [])
    if len(fcs) != 2:
        raise ValueError("checkpoint: expected fc6 and fc7 in recog_base, found %d weighted modules" % len(fcs))
    for name, m in zip(("fc6", "fc7"), fcs):
        w[name + ".weight"] = np.asarray(m.fields["weight"], np.float32)
        w[name + ".bias"] = np.asarray(m.fields["bias"], np.float32).reshape(-1)
    for name, key in (("obj", "objectness_branch"), ("boxreg", "box_reg_branch")):
        m = nets[key]
        w[name + ".weight"] = np.asarray(m.fields["weight"], np.float32)
        w[name + ".bias"] = np.asarray(m.fields["bias"], np.float32).reshape(-1)
    lm = nets["language_model"]
    enc = _walk_modules(_field(lm, "image_encoder"), [])[0]
    w["lm.encoder.weight"] = np.asarray(enc.fields["weight"], np.float32)
    w["lm.encoder.bias"] = np.asarray(enc.fields["bias"], np.float32).reshape(-1)
    w["lm.lookup.weight"] = np.asarray(_field(lm, "lookup_table").fields["weight"], np.float32)
    rnn = _walk_modules(_field(lm, "rnn"), [])
    lstm = [m for m in rnn if m.classname.endswith("LSTM")]
    lin = [m for m in rnn if not m.classname.endswith("LSTM")]
    if len(lstm) != 1 or len(lin) != 1:
        raise ValueError("checkpoint: expected one nn.LSTM and one nn.Linear in language_model.rnn")
    w["lm.lstm.weight"] = np.asarray(lstm[0].fields["weight"], np.float32)           # torch-rnn: ((D+H), 4H), gates [i|f|o|g]
    w["lm.lstm.bias"] = np.asarray(lstm[0].fields["bias"], np.float32).reshape(-1)
    w["lm.vocab.weight"] = np.asarray(lin[0].fields["weight"], np.float32)
    w["lm.vocab.bias"] = np.asarray(lin[0].fields["bias"], np.float32).reshape(-1)
    lf = lm.fields if isinstance(lm, T7Object) else lm
    opt = model.fields.get("opt") if isinstance(model, T7Object) else None
    info = {"vocab_size": int(lf.get("vocab_size", w["lm.vocab.weight"].shape[0] - 1)), "seq_length": lf.get("seq_length"),
            "idx_to_token": lf.get("idx_to_token"), "opt": opt}
    info.update(_model_state(nets))
    return w, info


def _walk_all(obj, out, seen=None):
    """Every T7Object of a module tree (tables and nn.Sequential-like `modules` lists included), each once."""
    seen = seen if seen is not None else set()
    if id(obj) in seen:
        return out
    seen.add(id(obj))
    if isinstance(obj, T7Object):
        out.append(obj)
        _walk_all(obj.fields, out, seen)
    elif isinstance(obj, dict):
        for v in obj.values():
            _walk_all(v, out, seen)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            _walk_all(v, out, seen)
    return out


def _model_state(nets):
    """Non-weight state a checkpoint carries and the context must honour instead of its defaults:
      anchors (2, k) and the receptive-field centres x0, y0, sx, sy of nn.MakeAnchors (MakeAnchors.lua:31-38, LocalizationLayer.lua:653-655),
      ceil_mode of the VGG nn.SpatialMaxPooling layers (DenseCapModel.lua:61-76 sets it on the loaded VGG-16).
    Missing fields (e.g. a tree written by densecap_checkpoint_tree without them) leave the entry out -> library defaults."""
    st = {}
    mods = _walk_all(nets.get("localization_layer"), [])
    for m in mods:
        if m.classname.endswith("MakeAnchors"):
            f = m.fields
            if "anchors" in f and f["anchors"] is not None:
                st["anchors"] = np.asarray(f["anchors"], np.float32).reshape(2, -1)
            if all(k in f for k in ("x0", "y0", "sx", "sy")):
                st["field_centers"] = tuple(float(f[k]) for k in ("x0", "y0", "sx", "sy"))
            break
    pools = [m for m in _walk_all(nets.get("conv_net1"), []) + _walk_all(nets.get("conv_net2"), []) if m.classname.endswith("SpatialMaxPooling")]
    modes = {bool(m.fields.get("ceil_mode", False)) for m in pools}
    if len(modes) == 1:
        st["pool_ceil"] = modes.pop()
    elif len(modes) > 1:
        raise ValueError("checkpoint: the VGG pooling layers mix ceil and floor mode, which the trunk kernels do not support")
    return st


# ---------------------------------------------------------------------------------------------------------
# the inverse direction: weights (dc_load_weights names) -> the object tree train.lua:157-185 saves.  Used by the tests and to
# export synthetic / converted weights in the reference's container (`save(densecap_checkpoint_tree(dims, w), "model.t7")`).
# ---------------------------------------------------------------------------------------------------------
def _seq(*mods):
    return T7Object("nn.Sequential", {"modules": list(mods)})


def _conv(w, b, mm=False):
    cout = w.shape[0]
    return T7Object("nn.SpatialConvolutionMM" if mm else "nn.SpatialConvolution",
                       {"weight": w.reshape(cout, -1) if mm else w, "bias": b, "kW": w.shape[3], "kH": w.shape[2]})


def _lin(w, b):
    return T7Object("nn.Linear", {"weight": w, "bias": b})


def _relu():
    return T7Object("nn.ReLU", {"inplace": True})


def densecap_checkpoint_tree(dims, w):
    """The module tree of DenseCapModel.lua:40-125 / LocalizationLayer.lua:609-690 / LanguageModel.lua:26-71, weights from `w`."""
    names = VGG16_CONV_NAMES
    vgg = []
    for i, n in enumerate(names):
        vgg += [_conv(w[n + ".weight"], w[n + ".bias"], mm=(i % 2 == 0)), _relu()]
        if n in ("conv1_2", "conv2_2", "conv3_3", "conv4_3"):
            vgg.append(T7Object("nn.SpatialMaxPooling", {"kW": 2, "kH": 2, "ceil_mode": True}))
    conv_net1 = _seq(*vgg[:10])
    conv_net2 = _seq(*vgg[10:])
    box_branch = _seq(_conv(w["rpn.box.weight"], w["rpn.box.bias"]), T7Object("nn.RegularizeLayer", {"w": 0}),
                     T7Object("nn.ConcatTable", {"modules": [_seq(T7Object("nn.MakeAnchors", {}), T7Object("nn.ReshapeBoxFeatures", {})),
                                                                T7Object("nn.ReshapeBoxFeatures", {})]}),
                     T7Object("nn.ConcatTable", {"modules": [T7Object("nn.ApplyBoxTransform", {}), T7Object("nn.Identity", {})]}))
    rpn_branch = _seq(_conv(w["rpn.score.weight"], w["rpn.score.bias"]), T7Object("nn.ReshapeBoxFeatures", {}))
    rpn = _seq(_conv(w["rpn.conv.weight"], w["rpn.conv.bias"]), _relu(), T7Object("nn.ConcatTable", {"modules": [box_branch, rpn_branch]}),
              T7Object("nn.FlattenTable", {}))
    loc = T7Object("nn.LocalizationLayer", {"nets": {"rpn": rpn, "roi_pooling": T7Object("nn.BilinearRoiPooling", {})}, "opt": {"std": 0.01}})
    recog = _seq(T7Object("nn.View", {}), _lin(w["fc6.weight"], w["fc6.bias"]), _relu(), T7Object("nn.Dropout", {"p": 0.5}),
                _lin(w["fc7.weight"], w["fc7.bias"]), _relu(), T7Object("nn.Dropout", {"p": 0.5}))
    lookup = T7Object("nn.LookupTable", {"weight": w["lm.lookup.weight"]})
    enc = _seq(_lin(w["lm.encoder.weight"], w["lm.encoder.bias"]), _relu(), T7Object("nn.View", {}))
    rnn = _seq(T7Object("nn.LSTM", {"weight": w["lm.lstm.weight"], "bias": w["lm.lstm.bias"]}), T7Object("nn.View", {}),
              _lin(w["lm.vocab.weight"], w["lm.vocab.bias"]), T7Object("nn.View", {}))
    lm = T7Object("nn.LanguageModel", {"vocab_size": dims.vocab_size, "seq_length": dims.seq_length, "image_encoder": enc, "lookup_table": lookup,
                                          "rnn": rnn, "idx_to_token": {i + 1: "w%d" % i for i in range(dims.vocab_size)},
                                          "net": _seq(T7Object("nn.ParallelTable", {"modules": [enc, lookup]}), rnn)})   # shared references
    model = T7Object("DenseCapModel", {"nets": {"conv_net1": conv_net1, "conv_net2": conv_net2, "localization_layer": loc, "recog_base": recog,
                                                   "objectness_branch": _lin(w["obj.weight"], w["obj.bias"]),
                                                   "box_reg_branch": _lin(w["boxreg.weight"], w["boxreg.bias"]), "language_model": lm},
                                          "opt": {"cnn_name": "vgg-16", "rpn_num_filters": dims.rpn_filters}})
    return {"opt": {"checkpoint_path": "x.t7"}, "iter": 1000, "model": model}


