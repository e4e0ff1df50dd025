"""Import the reference's scenario inputs (data/<name>/) into this repo's data/ directory.

Inputs are data, not code: YAML is re-dumped in canonical form (comments dropped, values
untouched); adj_matrix.npy / edge_distances.pkl / node_positions.json are copied byte for
byte.  Run once in the build container (needs /root/reference).
"""
import os, shutil, sys, yaml

SRC = "/root/reference/data"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data")

def main():
    for name in sorted(os.listdir(SRC)):
        d = os.path.join(SRC, name)
        if not os.path.isdir(d) or not os.path.exists(os.path.join(d, "sim_params.yaml")):
            continue
        out = os.path.join(DST, name)
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(d, "sim_params.yaml")) as f:
            cfg = yaml.safe_load(f)
        with open(os.path.join(out, "sim_params.yaml"), "w") as f:
            f.write(f"# scenario '{name}': canonical dump of the PedNStream input of the same name\n")
            yaml.safe_dump(cfg, f, sort_keys=False, default_flow_style=None, width=100)
        for extra in ("adj_matrix.npy", "edge_distances.pkl", "node_positions.json"):
            p = os.path.join(d, extra)
            if os.path.exists(p):
                shutil.copyfile(p, os.path.join(out, extra))
        print("imported", name, sorted(os.listdir(out)))

if __name__ == "__main__":
    main()
