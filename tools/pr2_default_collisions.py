"""The "always in collision" rule of the MoveIt Setup Assistant (moveit_setup_assistant/src/tools/
compute_default_collisions.cpp:474-547: a link pair that collides in >= 95 % of the random states is disabled) applied
to the synthetic PR2-like tree with the oracle's sphere decomposition, on the CPU.  Prints the pairs
robot_description.pr2_like() then lists as disabled next to the adjacent ones, like the real PR2's SRDF does; without
them every random state self-collides (grandparent / grandchild cylinders overlap by construction)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main(n=1000, fraction=0.95):
    import oracle_py
    from moveit_b200 import robot_description as rd
    from moveit_b200 import scenarios
    from common import OracleSetup
    desc = rd.pr2_like()
    o = OracleSetup(oracle_py, desc, scenarios.FIELD_DEFAULT, [])
    link_of, radius = [], []
    for i, s in enumerate(o.model_desc["spheres"]):
        for row in s:
            link_of.append(i)
            radius.append(row[3])
    link_of, radius = np.array(link_of), np.array(radius)
    q = desc.sample_states(n, 2024)
    L = desc.n_links
    hits = np.zeros((L, L), dtype=np.int64)
    for k in range(n):
        c = np.asarray(o.env.sphere_centers(q[k]))
        d = np.linalg.norm(c[:, None, :] - c[None, :, :], axis=2)
        touch = d < radius[:, None] + radius[None, :]
        pair = np.zeros((L, L), dtype=bool)
        ii, jj = np.nonzero(touch)
        pair[link_of[ii], link_of[jj]] = True
        hits += pair
    always = []
    already = {tuple(sorted(p)) for p in desc.disabled_pairs}
    for i in range(L):
        for j in range(i + 1, L):
            if hits[i, j] >= fraction * n and tuple(sorted((desc.link_names[i], desc.link_names[j]))) not in already:
                always.append((desc.link_names[i], desc.link_names[j], hits[i, j] / n))
    for a, b, f in always:
        print('        ("%s", "%s"),  # %.3f' % (a, b, f))
    print("# %d pairs always in collision" % len(always))


if __name__ == "__main__":
    main()
