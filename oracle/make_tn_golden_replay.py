"""Recipe replay shared by oracle/make_tn_golden.py and tests/test_tn.py
(kept free of any /root/reference import so it also runs on the GPU box)."""
import numpy as np


def replay(TN, recipe):
    """Recipe ops: ["dense", shape] ["copy", degree] ["input", shape]
    ["conn", n1, n2, leg1, leg2] ["rm", node, leg]."""
    tn, nodes = TN(), []
    for op in recipe:
        kind = op[0]
        if kind == "dense":
            nodes.append(tn.add_dense_node(np.ones(tuple(op[1]))))
        elif kind == "copy":
            nodes.append(tn.add_copy_node(op[1]))
        elif kind == "input":
            nodes.append(tn.add_input_node(tuple(op[1])))
        elif kind == "conn":
            tn.connect_nodes(nodes[op[1]], nodes[op[2]], op[3], op[4])
        elif kind == "rm":
            tn.remove_edge(nodes[op[1]][op[2]])
    return tn
