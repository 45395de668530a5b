"""Dictionary trie of the reference's constrained decode (TEST INFRASTRUCTURE, see oracle/__init__.py).

`load_dictionary` restates loadDictionary (src/utils/utils.lua:177-218): a hash-of-hashes keyed by vocabulary id, rooted
at trie[2] (the GO symbol); every word ends in an EOS (3) child; with `allow_digit_prefix` the root loops to itself on
EOS and on every digit.  Nodes here are Python dicts {vocab_id: node}; key 0 (not a vocabulary id) holds the order in
which the node was created, which `flatten` uses to number the nodes (root = 0) of the (num_nodes, V + 1) child table
the library consumes (-1 = no child, column = 1-based vocabulary id)."""
import numpy as np


def load_dictionary(words, allow_digit_prefix=False):
    count = [0]

    def new():
        n = {0: count[0]}
        count[0] += 1
        return n
    root = new()
    for w in words:
        s = w.strip()
        node = root
        if allow_digit_prefix:                      # utils.lua:195-201
            node[3] = root
            for l in range(48, 58):
                node[l - 48 + 3 + 1] = root
        for ch in s:                                # utils.lua:202-214
            l = ord(ch)
            vid = l - 97 + 13 + 1 if l > 96 else l - 48 + 3 + 1
            if vid not in node:
                node[vid] = new()
            node = node[vid]
        if 3 not in node:                           # utils.lua:215-217
            node[3] = new()
    return root


def flatten(root, V=39):
    seen, stack = {}, [root]
    while stack:
        n = stack.pop()
        if n[0] in seen:
            continue
        seen[n[0]] = n
        stack.extend(c for k, c in n.items() if k != 0)
    table = -np.ones((len(seen), V + 1), np.int32)
    for i, node in seen.items():
        for vid, child in node.items():
            if vid != 0:
                table[i, vid] = child[0]
    return table
