"""Drop-in for /root/reference/node_classification/scalers.py (PNA-style degree scalers).

Same names and signatures (`SCALERS[name](input, add_all, num_aggregators, avg_d=None)`),
same arithmetic, minus the hard-coded device 'cuda:2' (scalers.py:12-61): everything
follows `input.device`.  `add_all` may be a list of neighbour arrays or -- as the reference's
own caller passes it (layers.py:856, Q7) -- the sparse adjacency, in which case every
"degree" is len(adj[i]) == N.  The scale tensor is tiled only for num_aggregators in 1..4
(scalers.py:33-40, 51-58); more aggregators fail with torch's broadcast error, as upstream.
"""
import torch


def _all_degrees(add_all, device):
    if isinstance(add_all, torch.Tensor):
        n = add_all.shape[0]
        width = add_all.shape[1] if add_all.dim() > 1 else 1
        return torch.full((n,), width, dtype=torch.int64, device=device)     # len(adj[i]) == N for every row
    return torch.tensor([len(node_nei) for node_nei in add_all], device=device)


def avg_d_log(all_degrees):
    return torch.mean(torch.log(all_degrees + 1))


def avg_d_exp(all_degrees):
    return torch.mean(torch.exp(torch.div(1, all_degrees)) - 1)


def scale_identity(input, add_all, num_aggregators, avg_d=None):
    return input


def _tile(scale, num_aggregators):
    if num_aggregators in (2, 3, 4):
        scale = torch.cat((scale,) * num_aggregators, 0)
    return scale


def scale_amplification(input, add_all, num_aggregators, avg_d=None):
    all_degrees = _all_degrees(add_all, input.device)
    scale = (torch.log(all_degrees + 1) / avg_d_log(all_degrees)).unsqueeze(-1)
    return torch.mul(_tile(scale, num_aggregators), input)


def scale_attenuation(input, add_all, num_aggregators, avg_d=None):
    all_degrees = _all_degrees(add_all, input.device)
    scale = (avg_d_log(all_degrees) / torch.log(all_degrees + 1)).unsqueeze(-1)
    return torch.mul(_tile(scale, num_aggregators), input)


SCALERS = {"identity": scale_identity, "amplification": scale_amplification, "attenuation": scale_attenuation}
