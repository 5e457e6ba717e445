#!/usr/bin/env python
"""Prints the node tree of a flattened scene (scenes/<name>.npz)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import actinon_b200 as acn
from collections import Counter
KN = ['COMPOUND', 'PLANE', 'SPHERE', 'SQUAROID', 'DSPHERE', 'TORUS', 'AND', 'OR', 'NEG', 'SCALE']
name = sys.argv[1]
maxind = int(sys.argv[2]) if len(sys.argv) > 2 else 12
flat = acn.scenes.load(name)
st = flat.struct
nodes = [st.nodes[i] for i in range(st.n_nodes)]
ch = [st.children[i] for i in range(st.n_children)]
print(name, 'n_nodes', len(nodes), Counter(KN[n.kind] for n in nodes))
def show(i, ind):
    n = nodes[i]
    k = KN[n.kind]
    env = f" env r={n.env_radius:.3g}" if n.has_envelope else ""
    if ind > maxind:
        print(' ' * ind + '...'); return
    if n.kind == 0:
        print(' ' * ind + f"[{i}] COMPOUND x{n.child1}{env}")
        for c in range(min(n.child1, 12)): show(ch[n.child0 + c], ind + 2)
    elif n.kind in (6, 7):
        print(' ' * ind + f"[{i}] {k}{env} mat={n.material} rough={n.surface_roughness}")
        show(n.child0, ind + 2); show(n.child1, ind + 2)
    elif n.kind in (8, 9):
        print(' ' * ind + f"[{i}] {k}{env}"); show(n.child0, ind + 2)
    else:
        print(' ' * ind + f"[{i}] {k}{env} mat={n.material} rough={n.surface_roughness} tail={list(n.tail)[:4]}")
show(st.light_root, 0); show(st.matter_root, 0)
p = flat.params
print('ds', p.direct_samples, 'ps', p.path_samples, 'depth', p.trace_depth, 'minI', p.trace_min_intensity, 'maxpath', p.max_path_length, p.image_width, p.image_height)
