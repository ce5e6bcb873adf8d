"""Network visualisation (reference: python/mxnet/visualization.py): textual summary of a Symbol."""
import json


def print_summary(symbol, shape=None, line_length=100):
    conf = json.loads(symbol.tojson())
    shapes = {}
    if shape is not None:
        vshape, _o, _d, _t = symbol.get_internals()._infer(shape, partial=True)
        shapes = vshape
    print("%-40s %-20s %s" % ("Layer (type)", "Op", "Inputs"))
    print("=" * line_length)
    total = 0
    for node in conf["nodes"]:
        if node["op"] == "null":
            shp = shapes.get(node["name"])
            if shp is not None and not node["name"].endswith(("data", "label")):
                n = 1
                for d in shp:
                    n *= d
                total += n
            continue
        print("%-40s %-20s %s" % (node["name"], node["op"], ",".join(conf["nodes"][i[0]]["name"] for i in node["inputs"])))
    print("=" * line_length)
    print("Total params: %d" % total)
    return total


def plot_network(symbol, **kwargs):
    raise RuntimeError("graphviz is not available in this environment; use print_summary")
