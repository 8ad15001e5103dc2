#!/usr/bin/env python3
"""Sample real 256-bit ORB descriptors from the reference's binary vocabulary into a small fixture.

Vocabulary/ORBvoc.bin layout (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1442-1478): header
{uint32 nb_nodes, uint32 size_node, uint32 k, uint32 L, ...} then 41-byte nodes
{int32 parent; uint8 desc[32]; float weight; uint8 is_leaf}.  The reference tree is absent on the GPU
box, so 4096 leaf descriptors (evenly strided) are committed as tests/golden/orbvoc_sample.npy.
"""
import os
import struct
import sys
import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
raw = open(os.path.join(ref, "Vocabulary", "ORBvoc.bin"), "rb").read()
nb_nodes, size_node = struct.unpack_from("<II", raw, 0)
assert size_node == 41, size_node
hdr = len(raw) - nb_nodes * size_node   # the root is not stored / header size: whatever precedes the node table
if hdr < 0:
    nb_nodes -= 1
    hdr = len(raw) - nb_nodes * size_node
nodes = np.frombuffer(raw, np.uint8, offset=hdr).reshape(nb_nodes, 41)
leaf = nodes[:, 40] == 1
desc = nodes[leaf][:, 4:36]
step = len(desc) // 4096
sample = np.ascontiguousarray(desc[::step][:4096])
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
np.save(os.path.join(root, "tests", "golden", "orbvoc_sample.npy"), sample)
print("nodes", nb_nodes, "header", hdr, "leaves", int(leaf.sum()), "sample", sample.shape,
      "mean popcount", np.unpackbits(sample, axis=1).sum(1).mean())
