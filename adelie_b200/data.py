"""Synthetic data generators with the semantics of ``adelie.data`` (reference: adelie/data.py:13-219):
``np.random.seed(seed)``; X ~ N(0,1) column-major; beta* ~ N(0,1) on a random (1-sparsity) support;
gaussian y = eta + ||beta*|| N(0,1)/sqrt(snr); binomial y ~ Bernoulli(sigmoid(eta/||beta*||))."""
import numpy as np

from . import glm as _glm


def _sample_y(glm, eta, beta, rho=0, snr=1):
    n, K = eta.shape
    is_multi = "multi" in glm
    if not is_multi and K > 1:
        eta = eta[:, 0][:, None]
        K = 1
    if "gaussian" in glm:
        signal_scale = np.sqrt(rho * np.sum(beta) ** 2 + (1 - rho) * np.sum(beta ** 2))
        noise_scale = signal_scale / np.sqrt(snr)
        y = eta + noise_scale * np.random.normal(0, 1, eta.shape)
        if is_multi:
            return _glm.multigaussian(y=y)
        return _glm.gaussian(y=y.ravel())
    if glm == "binomial":
        scale = np.sqrt(rho * np.sum(beta) ** 2 + (1 - rho) * np.sum(beta ** 2))
        eta = eta / max(scale, 1e-300)
        mu = 1 / (1 + np.exp(-eta))
        y = np.random.binomial(1, mu).astype(np.float64)
        return _glm.binomial(y=y.ravel())
    if glm == "poisson":                                   # mu = exp(eta / scale), y ~ Poisson(mu)  (adelie/data.py:63-81)
        scale = np.sqrt(rho * np.sum(beta) ** 2 + (1 - rho) * np.sum(beta ** 2))
        mu = np.exp(eta.ravel() / max(scale, 1e-300))
        return _glm.poisson(y=np.random.poisson(mu).astype(np.float64))
    if glm == "cox":
        scale = np.sqrt(rho * np.sum(beta) ** 2 + (1 - rho) * np.sum(beta ** 2))
        eta = (eta / max(scale, 1e-300)).ravel()
        s = np.random.exponential(1, n)
        t = s + 1 + np.random.exponential(np.exp(-eta))
        c = s + 1 + np.random.exponential(1, n)
        d = (t <= c).astype(np.float64)
        t = np.minimum(t, c)
        return _glm.cox(start=s, stop=t, status=d)
    raise RuntimeError(f"unsupported glm {glm}")


def dense(n, p, G, *, K=1, glm="gaussian", equal_groups=False, rho=0, sparsity=0.95, zero_penalty=0, snr=1, seed=0):
    """Dense dataset (semantics of adelie/data.py:84-219)."""
    assert n >= 1 and p >= 1 and G >= 1
    np.random.seed(seed)
    if equal_groups:
        groups = (p // G) * np.arange(G)
    else:
        groups = np.concatenate([[0], np.random.choice(np.arange(1, p), size=G - 1, replace=False)])
        groups = np.sort(groups).astype(int)
    group_sizes = np.concatenate([groups, [p]], dtype=int)
    group_sizes = group_sizes[1:] - group_sizes[:-1]
    penalty = np.sqrt(group_sizes)
    penalty[np.random.choice(G, int(zero_penalty * G), replace=False)] = 0
    penalty /= np.linalg.norm(penalty) / np.sqrt(p)
    X = np.random.normal(0, 1, (n, p))
    Z = np.random.normal(0, 1, n)
    X = np.sqrt(rho) * Z[:, None] + np.sqrt(1 - rho) * X
    X = np.asfortranarray(X)
    beta = np.random.normal(0, 1, (p, K))
    beta_zero_indices = np.random.choice(p, int(sparsity * p), replace=False)
    beta_nnz_indices = np.array(list(set(np.arange(p)) - set(beta_zero_indices)))
    X_sub = X[:, beta_nnz_indices]
    beta_sub = beta[beta_nnz_indices]
    eta = X_sub @ beta_sub
    glm_obj = _sample_y(glm=glm, eta=eta, beta=beta_sub, rho=rho, snr=snr)
    return {"X": X, "glm": glm_obj, "groups": groups, "group_sizes": group_sizes, "penalty": penalty}


def snp_unphased(n, p, *, K=1, glm="gaussian", sparsity=0.95, missing_ratio=0.1, one_ratio=0.25, two_ratio=0.05, zero_penalty=0, snr=1, seed=0):
    """SNP unphased dataset (semantics and RNG call order of adelie/data.py:222-359): int8 calldata with ``one_ratio`` ones,
    ``two_ratio`` twos, the response drawn from the complete matrix, then ``missing_ratio`` of the entries masked as -9."""
    assert n >= 1 and p >= 1 and snr > 0 and seed >= 0
    for r in (sparsity, missing_ratio, one_ratio, two_ratio, zero_penalty):
        assert 0 <= r <= 1
    np.random.seed(seed)
    nz_ratio = one_ratio + two_ratio
    n_nz = int(nz_ratio * n * p)
    where = np.random.permutation(np.random.choice(n * p, n_nz, replace=False))
    n_ones = int(one_ratio / nz_ratio * n_nz)
    X = np.zeros((n, p), dtype=np.int8)
    X.ravel()[where[:n_ones]] = 1
    X.ravel()[where[n_ones:]] = 2
    groups = np.arange(p)
    group_sizes = np.ones(p, dtype=int)
    penalty = np.sqrt(group_sizes)
    penalty[np.random.choice(p, int(zero_penalty * p), replace=False)] = 0
    penalty /= np.linalg.norm(penalty) / np.sqrt(p)
    beta = np.random.normal(0, 1, (p, K))
    support = np.random.choice(p, int((1 - sparsity) * p), replace=False)
    beta_sub = beta[support]
    eta = X[:, support] @ beta_sub
    glm_obj = _sample_y(glm=glm, eta=eta, beta=beta_sub, snr=snr)
    X.ravel()[np.random.choice(n * p, int(missing_ratio * n * p), replace=False)] = -9
    return {"X": np.asfortranarray(X), "glm": glm_obj, "groups": groups, "group_sizes": group_sizes, "penalty": penalty}


def snp_phased_ancestry(n, s, A, *, K=1, glm="gaussian", sparsity=0.95, one_ratio=0.25, two_ratio=0.05, zero_penalty=0, snr=1, seed=0):
    """SNP phased, ancestry dataset (semantics of adelie/data.py ``snp_phased_ancestry``): int8 ``X`` (n, 2 s) with mutation indicators
    such that a fraction ``one_ratio`` / ``two_ratio`` of the (individual, SNP) pairs carries 1 / 2 mutations, uniform random ancestry labels
    ``ancestries`` (n, 2 s) in [0, A), groups = one group of A ancestry columns per SNP, response drawn from the dense (n, s A) matrix."""
    assert n >= 1 and s >= 1 and A >= 1 and snr > 0 and seed >= 0
    np.random.seed(seed)
    nz_ratio = one_ratio + two_ratio
    n_nz = int(nz_ratio * n * s)
    where = np.random.permutation(np.random.choice(n * s, n_nz, replace=False))
    n_ones = int(one_ratio / nz_ratio * n_nz)
    X = np.zeros((n, s), dtype=np.int8)
    X.ravel()[where[:n_ones]] = 1
    X.ravel()[where[n_ones:]] = 2
    caldata = np.zeros((n, 2 * s), dtype=np.int8)
    caldata[:, ::2] = X >= 1
    caldata[:, 1::2] = X >= 2
    ancestries = np.zeros((n, 2 * s), dtype=np.int8)
    ancestries[:, ::2] = np.random.choice(A, (n, s), replace=True)
    ancestries[:, 1::2] = np.random.choice(A, (n, s), replace=True)
    groups = A * np.arange(s)
    group_sizes = np.full(s, A)
    penalty = np.sqrt(group_sizes.astype(float))
    penalty[np.random.choice(s, int(zero_penalty * s), replace=False)] = 0
    penalty /= np.linalg.norm(penalty) / np.sqrt(s * A)
    dense = np.zeros((n, s * A), dtype=np.int8)
    for k in range(2):
        rows_i, snp_j = np.nonzero(caldata[:, k::2])
        np.add.at(dense, (rows_i, snp_j * A + ancestries[:, k::2][rows_i, snp_j]), 1)
    beta = np.random.normal(0, 1, (s * A, K))
    support = np.random.choice(s * A, int((1 - sparsity) * s * A), replace=False)
    beta_sub = beta[support]
    eta = dense[:, support].astype(float) @ beta_sub
    glm_obj = _sample_y(glm=glm, eta=eta, beta=beta_sub, snr=snr)
    return {"X": np.asfortranarray(caldata), "ancestries": np.asfortranarray(ancestries), "glm": glm_obj, "groups": groups,
            "group_sizes": group_sizes, "penalty": penalty, "dense": dense}
