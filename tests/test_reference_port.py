"""CPU: oracle/reference_port.py (the dense torch restatement timed as the CPU baseline) against the
real reference's golden runs."""
import numpy as np
import pytest
import torch

from oracle import fk as OFK
from oracle import planner as OP
from oracle.reference_port import ReferencePort

from helpers import GOLDEN, load, n_iters, rel


@pytest.mark.parametrize("name", GOLDEN)
def test_port_reproduces_reference(name):
    g = load(name)
    spec = OP.spec_from_golden(g)
    dtype = torch.float32 if spec['dtype'] == 'float32' else torch.float64
    port = ReferencePort(spec, dtype=dtype, fk=OFK.fk_all_links_torch() if ('spheres' in spec or 'self_margin' in spec) else None)
    assert rel(port.Sigma_inv.numpy(), g['Sigma_inv']) < (1e-6 if dtype == torch.float32 else 1e-15)
    tol = 2e-3 if dtype == torch.float32 else 1e-9
    for it in range(n_iters(g)):
        pre = f'it{it}_'
        port.set_mean(torch.tensor(g[pre + 'means_pre']))
        if dtype == torch.float64:
            assert rel(port.L[0].numpy(), g['L']) < 1e-9
        samples, costs, w, grad = port.iterate(torch.tensor(g[pre + 'eps']))
        assert rel(samples.numpy(), g[pre + 'samples']) < tol
        assert rel(costs.numpy(), g[pre + 'costs']) < tol
        if dtype == torch.float64:
            assert np.abs(w.numpy() - g[pre + 'weights']).max() < 1e-7
            assert rel(port.means.numpy(), g[pre + 'means_post']) < 1e-9
