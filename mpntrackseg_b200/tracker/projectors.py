"""Rounding of the edge predictions of the sequence graph into a feasible flow, on the GPU.
reference: src/mot_neural_solver/tracker/projectors.py (GreedyProjector :11-67, ExactProjector :69-113)."""
from .. import ops


class GreedyProjector:
    """Greedy rounding (https://arxiv.org/pdf/1912.07515.pdf, appendix B.1): threshold at 0.5, then every node whose
    outgoing (then incoming) flow exceeds 1 keeps its highest-scoring active edge.  ``full_graph.graph_obj`` is the
    undirected, pruned sequence graph (one entry per pair, row = earlier node).  reference: tracker/projectors.py:11-67"""

    def __init__(self, full_graph):
        self.final_graph = full_graph.graph_obj
        self.num_nodes = full_graph.graph_obj.num_nodes

    def project(self):
        round_preds, self.constr_satisf_rate = ops.greedy_project(self.final_graph.edge_index, self.final_graph.edge_preds,
                                                                  self.num_nodes)
        self.final_graph.edge_preds = round_preds


class ExactProjector:
    """Min-cost-flow rounding through a linear program (reference: tracker/projectors.py:69-113, PuLP / Gurobi on the
    host).  The LP solver is a third-party host library outside the hot path (SURVEY.md f2 keeps it a CPU step); it is
    not part of this package."""

    def __init__(self, full_graph, solver_backend='pulp'):
        self.final_graph = full_graph.graph_obj
        self.solver_backend = solver_backend

    def project(self):
        raise NotImplementedError("rounding_method 'exact' needs the reference's PuLP linear program on the host; "
                                  "use rounding_method 'greedy' (GreedyProjector) on the GPU")
