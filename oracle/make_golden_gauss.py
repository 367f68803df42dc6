"""Generate tests/golden/gauss_primitives.npz from the UNMODIFIED reference classes in
/root/reference/rbnet/multivariate_normal.py (PairwiseProduct, Product, ApproximateMixture).  Build container only.
Run:  python -m oracle.make_golden_gauss"""
import os

import numpy as np
import torch

from oracle.reference_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    torch.set_default_dtype(torch.float64)
    mvn = load_reference()["multivariate_normal"]
    rng = np.random.default_rng(0)      # the reference's own tests seed default_rng(0) (test_multivariate_normal.py:15)
    store = {}
    k = 0
    for D in (1, 2, 4, 8):
        for batch in ((), (3,), (2, 5)):
            def spd():
                a = rng.uniform(-1, 1, batch + (D, D))
                return a @ np.swapaxes(a, -1, -2) + 0.1 * np.eye(D)
            m1, m2 = rng.standard_normal(batch + (D,)), rng.standard_normal(batch + (D,))
            c1, c2 = spd(), spd()
            pp = mvn.PairwiseProduct(mean1=torch.tensor(m1), mean2=torch.tensor(m2), cov1=torch.tensor(c1), cov2=torch.tensor(c2))
            pr = mvn.Product(means=torch.tensor(np.stack([m1, m2])), covariances=torch.tensor(np.stack([c1, c2])))
            ln, mean, cov = pr.product()
            store[f"pp{k}/m1"], store[f"pp{k}/m2"], store[f"pp{k}/c1"], store[f"pp{k}/c2"] = m1, m2, c1, c2
            store[f"pp{k}/log_norm"] = pp.log_norm.numpy()
            store[f"pp{k}/mean"] = pp.mean.numpy()
            store[f"pp{k}/cov"] = pp.cov.numpy()
            store[f"pp{k}/product_log_norm"] = ln.numpy()
            store[f"pp{k}/product_mean"] = mean.numpy()
            store[f"pp{k}/product_cov"] = cov.numpy()
            n = 5
            means = rng.standard_normal((n,) + batch + (D,))
            lw = rng.standard_normal((n,) + batch)
            covs = np.stack([spd() for _ in range(n)])
            am = mvn.ApproximateMixture(means=torch.tensor(means), log_weights=torch.tensor(lw), covariances=torch.tensor(covs))
            store[f"am{k}/means"], store[f"am{k}/lw"], store[f"am{k}/covs"] = means, lw, covs
            store[f"am{k}/log_norm"] = am.log_norm.numpy()
            store[f"am{k}/mean"] = am.mean.numpy()
            store[f"am{k}/cov"] = am.covariance.numpy()
            k += 1
    store["n"] = np.asarray(k)
    np.savez_compressed(os.path.join(OUT, "gauss_primitives.npz"), **store)
    print("wrote", k, "cases")


if __name__ == "__main__":
    main()
