#!/usr/bin/env python
"""Exports the weights of the reference's shipped networks into safemotionsrisk_b200/assets/networks_<scene>.npz.

Run in the build container (it reads /root/reference; the result is committed, the GPU box never sees the reference):
    python tools/export_networks.py [/root/reference]

Sources (section 4):
  * backup policies  trained_networks/backup_networks/<scene>/checkpoint/checkpoint
      RLlib checkpoint = pickle of {'worker': pickle({'state': {'default_policy': {name: ndarray}}, ...})}; the ray
      classes inside are replaced by a stub while unpickling.  Model: keras_fcnet_last_layer_activation.py:86-136
      (fc_1, fc_2 swish -> fc_out tanh, 2 * n_joints outputs; the first n_joints are the action mean).
  * risk networks    trained_networks/risk_networks/state_action/<scene>/  (Keras SavedModel)
      Only the variables are needed: variables/variables.index is a TensorFlow tensor-bundle index (a LevelDB-format
      table of BundleEntryProto records), variables/variables.data-00000-of-00001 holds the raw tensors.  Both are
      parsed here without TensorFlow.  Model: Dense 512 / 256 / 128 selu + Dense 1 sigmoid on [observation, action]
      (train_risk_network.py; layer list in keras_metadata.pb).
"""
import io
import json
import os
import pickle
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------- RLlib checkpoint
class _Stub:
    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"state": state})


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("ray"):
            return _Stub
        return super().find_class(module, name)


def read_rllib_policy(path):
    outer = _Unpickler(open(path, "rb")).load()
    worker = _Unpickler(io.BytesIO(outer["worker"])).load()
    state = worker["state"]["default_policy"]
    return {k.split("default_policy/")[-1]: np.asarray(v) for k, v in state.items() if hasattr(v, "shape")}


# ---------------------------------------------------------------------------------------------- TF tensor bundle
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(block):
    """(key, value) pairs of one LevelDB table block (prefix-compressed keys, restart array at the end)."""
    n_restarts = struct.unpack("<I", block[-4:])[0]
    limit = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b""
    while pos < limit:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _read_block(data, offset, size):
    if data[offset + size] != 0:
        raise NotImplementedError("compressed table block (type {})".format(data[offset + size]))
    return data[offset:offset + size]


def _parse_proto(buf):
    """Minimal protobuf reader: {field: [values]} with varints as int and length-delimited fields as bytes."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _varint(buf, pos)
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wire == 5:
            val = struct.unpack("<I", buf[pos:pos + 4])[0]
            pos += 4
        elif wire == 1:
            val = struct.unpack("<Q", buf[pos:pos + 8])[0]
            pos += 8
        else:
            raise NotImplementedError("wire type {}".format(wire))
        out.setdefault(field, []).append(val)
    return out


def read_tf_bundle(prefix):
    """{variable name: ndarray} of a TensorFlow tensor bundle (float32 tensors only)."""
    index = open(prefix + ".index", "rb").read()
    data = open(prefix + ".data-00000-of-00001", "rb").read()
    footer = index[-48:]
    assert footer[-8:] == struct.pack("<Q", 0xDB4775248B80FB57), "not a LevelDB-format table"
    pos = 0
    _, pos = _varint(footer, pos)      # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    tensors = {}
    for _, handle in _block_entries(_read_block(index, idx_off, idx_size)):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, value in _block_entries(_read_block(index, off, size)):
            if key == b"":
                continue  # BundleHeaderProto
            entry = _parse_proto(value)
            dtype = entry.get(1, [0])[0]
            shape = []
            if 2 in entry:
                for dim in _parse_proto(entry[2][0]).get(2, []):
                    shape.append(_parse_proto(dim).get(1, [0])[0])
            offset, size = entry.get(4, [0])[0], entry.get(5, [0])[0]
            if dtype != 1:  # DT_FLOAT
                continue
            arr = np.frombuffer(data[offset:offset + size], dtype="<f4").reshape(shape).copy()
            tensors[key.decode()] = arr
    return tensors


def risk_layers(tensors):
    """Dense kernels / biases of the risk network in layer order (layer_with_weights-<i>/{kernel,bias})."""
    layers = []
    i = 0
    while True:
        k = "layer_with_weights-{}/kernel/.ATTRIBUTES/VARIABLE_VALUE".format(i)
        b = "layer_with_weights-{}/bias/.ATTRIBUTES/VARIABLE_VALUE".format(i)
        if k not in tensors:
            break
        layers.append((tensors[k], tensors[b]))
        i += 1
    return layers


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    nets = os.path.join(ref, "safemotions", "trained_networks")
    out_dir = os.path.join(ROOT, "safemotionsrisk_b200", "assets")
    for scene in ("space", "ball", "human"):
        out = {}
        if scene == "human":   # the policy that moves the human's arms (ctlp.py:4647-4762): 38 -> 256 -> 128 -> 16
            hp = read_rllib_policy(os.path.join(nets, "human_network", "checkpoint", "checkpoint"))
            for name in ("fc_1", "fc_2", "fc_out"):
                out["human/{}/kernel".format(name)] = hp[name + "/kernel"].astype(np.float32)
                out["human/{}/bias".format(name)] = hp[name + "/bias"].astype(np.float32)
            hcfg = json.load(open(os.path.join(nets, "human_network", "params.json")))
            out["human/log_std_range"] = np.asarray(hcfg["model"]["custom_model_config"]["log_std_range"], np.float32)
        pol = read_rllib_policy(os.path.join(nets, "backup_networks", scene, "checkpoint", "checkpoint"))
        for name in ("fc_1", "fc_2", "fc_out"):
            out["backup/{}/kernel".format(name)] = pol[name + "/kernel"].astype(np.float32)
            out["backup/{}/bias".format(name)] = pol[name + "/bias"].astype(np.float32)
        risk_dir = os.path.join(nets, "risk_networks", "state_action", scene)
        layers = risk_layers(read_tf_bundle(os.path.join(risk_dir, "variables", "variables")))
        assert len(layers) == 4, "expected Dense 512/256/128/1, found {} layers".format(len(layers))
        for i, (k, b) in enumerate(layers):
            out["risk/dense_{}/kernel".format(i)] = k
            out["risk/dense_{}/bias".format(i)] = b
        cfg = json.load(open(os.path.join(risk_dir, "risk_config.json")))
        out["risk/observation_size"] = np.int32(cfg["observation_size"])
        out["risk/action_size"] = np.int32(cfg["action_size"])
        path = os.path.join(out_dir, "networks_{}.npz".format(scene))
        np.savez_compressed(path, **out)
        print(path, {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    main()
