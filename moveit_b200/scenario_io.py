"""Writes a RobotDescription + world + request as the plain-text scenario the C++ host driver
(moveit_b200/host/host_driver.cpp) reads.  Numbers are printed with 17 significant digits so doubles round-trip."""
import numpy as np


def _f(v):
    return " ".join(repr(float(x)) for x in np.asarray(v, dtype=np.float64).reshape(-1))


def write_scenario(path, desc, field, objects, group="", use_acm=True, sphere_overrides=True, signed=False, tol=0.0,
                   n_single=0, n_contacts=0, max_contacts=1, max_contacts_per_pair=1, n_grad=0, mutate=False,
                   attached=(), extra_acm_always=()):
    """attached: [(name, link_name, [(type, dims, pose3x4)], [touch link names])];
    extra_acm_always: extra (a, b) name pairs set ALWAYS in the ACM (e.g. entries that name an attached body)."""
    lines = ["robot %s %d" % (desc.name, desc.n_links)]
    for i in range(desc.n_links):
        m = desc.mimic[i]
        lines.append("link %s %d %s %d %s %s %s %s %s" % (
            desc.link_names[i], desc.parent[i], desc.joint_names[i], desc.joint_type[i], _f(desc.axis[i]),
            _f(desc.origin[i]), m[0] if m else "-", repr(float(m[1])) if m else "1.0",
            repr(float(m[2])) if m else "0.0"))
        lines.append("%d" % len(desc.shapes[i]))
        for (t, dims, pose) in desc.shapes[i]:
            d = list(dims) + [0.0] * (3 - len(dims))
            lines.append("%d %s %s" % (t, _f(d), _f(pose)))
        name = desc.link_names[i]
        if sphere_overrides and name in desc.spheres_override and desc.shapes[i]:
            sp = desc.spheres_override[name]
            lines.append("%d %s" % (len(sp), _f(sp)))
        else:
            lines.append("-1")
    lines.append("groups %d" % len(desc.groups))
    for g, joints in desc.groups.items():
        lines.append("%s %d %s" % (g, len(joints), " ".join(joints)))
    pairs = list(desc.disabled_pairs) + list(extra_acm_always)
    lines.append("acm %d %d" % (1 if use_acm else 0, len(pairs)))
    for a, b in pairs:
        lines.append("%s %s" % (a, b))
    lines.append("field %s %s %r %r %d %r" % (_f(field["size"]), _f(field["center"]), float(field["resolution"]),
                                              float(field["max_distance"]), 1 if signed else 0, float(tol)))
    lines.append("objects %d" % len(objects))
    for k, (t, dims, pose) in enumerate(objects):
        d = list(dims) + [0.0] * (3 - len(dims))
        lines.append("obj%d %d %s %s" % (k, t, _f(d), _f(pose)))
    lines.append("attached %d" % len(attached))
    for name, link, shapes, touch in attached:
        lines.append("%s %s %d" % (name, link, len(shapes)))
        for (t, dims, pose) in shapes:
            d = list(dims) + [0.0] * (3 - len(dims))
            lines.append("%d %s %s" % (t, _f(d), _f(pose)))
        lines.append("%d %s" % (len(touch), " ".join(touch)))
    lines.append("request %s %d %d %d %d %d %d" % (group if group else "-", n_single, n_contacts, max_contacts,
                                                   max_contacts_per_pair, n_grad, 1 if mutate else 0))
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
