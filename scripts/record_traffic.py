"""Turns an ncu metrics log of scripts/ncu_target.py (dram__bytes_read.sum / dram__bytes_write.sum of pdl_main_kernel<true>) into
the profiles/traffic.json entry bench.py reports as roofline.traffic, stamped with the hash of the kernel it was captured for.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pdl_main_kernel \\
        --log-file gpurun_out/traffic_heat.log python scripts/ncu_target.py heat Tanh 1048576
    python scripts/record_traffic.py gpurun_out/traffic_heat.log heat tanh 1048576
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    log, w, act, B = sys.argv[1], sys.argv[2], sys.argv[3].lower(), int(sys.argv[4])
    import bench
    import pde_learn_b200 as P
    from pde_learn_b200 import _lib as L
    from pde_learn_b200 import configs, engine
    txt = open(log).read()
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = {}
    # the LAST launch in the log is the measured one (the earlier ones warm the plan up)
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.findall(re.escape(name) + r"\s+(\w+)\s+([0-9.,]+)", txt)
        if not m:
            raise SystemExit("no %s in %s" % (name, log))
        u, v = m[-1]
        vals[name] = float(v.replace(",", "")) * unit[u]
    d, bounds, derivs, lhs, rhs = configs.config(w)
    U = P.Network([d] + [40] * 5 + [1], "Rational" if act == "rational" else "Tanh")
    widths, a = engine.network_signature(U)
    encs, terms = engine.library_key(derivs, lhs, rhs)
    desc, keep = L.make_desc(d, widths, a, L.PDL_MODE_COLLOCATION, encs, [list(t) for t in terms])
    stamp = bench.kernel_stamp(L.plan_source(desc))
    path = os.path.join(ROOT, "profiles", "traffic.json")
    blob = json.load(open(path)) if os.path.exists(path) else {}
    blob["%s_%s" % (w, act)] = {"bytes": vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"], "read": vals["dram__bytes_read.sum"],
                                "write": vals["dram__bytes_write.sum"], "points_per_launch": B, "stamp": stamp,
                                "source": "ncu dram__bytes_read.sum + dram__bytes_write.sum of the last pdl_main_kernel<true> launch in %s"
                                          % os.path.basename(log)}
    json.dump(blob, open(path, "w"), indent=1)
    print(json.dumps(blob["%s_%s" % (w, act)]))


if __name__ == "__main__":
    main()
