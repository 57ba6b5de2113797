"""Flow-conservation statistics of a rounded solution on the GPU.
reference: src/mot_neural_solver/utils/evaluation.py:370-414 (compute_constr_satisfaction_rate)."""
from .. import ops


def compute_constr_satisfaction_rate(graph_obj, edges_out, undirected_edges=True, return_flow_vals=False):
    """Proportion of flow-conservation inequalities (sum of incoming, resp. outgoing, edge values <= 1 per node) that
    hold.  ``edges_out``: BINARISED edge values; ``undirected_edges``: every pair appears in both directions in
    ``graph_obj.edge_index`` (True) or once with row < col (False).  Returns the rate (Python float), and with
    ``return_flow_vals`` also (flow_in, flow_out) as [N] float tensors, like the reference."""
    rate, flow_in, flow_out = ops.constr_satisfaction(graph_obj.edge_index, edges_out.float(), graph_obj.num_nodes,
                                                      undirected_edges=undirected_edges)
    if not return_flow_vals:
        return rate
    return rate, flow_in, flow_out
