"""numpy restatement of the search's FID statistic (test oracle; see oracle/__init__.py).

Follows Evaluator.compute_statistics (evaluations/evaluator_v1.py:218-221) and
FIDStatistics.frechet_distance (evaluations/evaluator_v1.py:114-157; duplicate in
search_dynamic_unet_imagenet64_classifier_guidance_progressive.py:104-153). The only
deviation is API drift: scipy >= 1.16 dropped `sqrtm(..., disp=False)`, so the error
estimate it used to return (and the reference discards) is not requested.
"""
from __future__ import annotations

import warnings

import numpy as np
from scipy import linalg


def compute_statistics(activations: np.ndarray):
    mu = np.mean(activations, axis=0)
    sigma = np.cov(activations, rowvar=False)
    return mu, sigma


def frechet_distance(mu1, sigma1, mu2, sigma2, eps: float = 1e-6) -> float:
    mu1, mu2 = np.atleast_1d(mu1), np.atleast_1d(mu2)
    sigma1, sigma2 = np.atleast_2d(sigma1), np.atleast_2d(sigma2)
    assert mu1.shape == mu2.shape and sigma1.shape == sigma2.shape
    diff = mu1 - mu2
    covmean = linalg.sqrtm(sigma1.dot(sigma2))
    if not np.isfinite(covmean).all():
        warnings.warn("fid calculation produces singular product; adding %s to diagonal of cov estimates" % eps)
        offset = np.eye(sigma1.shape[0]) * eps
        covmean = linalg.sqrtm((sigma1 + offset).dot(sigma2 + offset))
    if np.iscomplexobj(covmean):
        if not np.allclose(np.diagonal(covmean).imag, 0, atol=1e-3):
            raise ValueError("Imaginary component {}".format(np.max(np.abs(covmean.imag))))
        covmean = covmean.real
    return float(diff.dot(diff) + np.trace(sigma1) + np.trace(sigma2) - 2 * np.trace(covmean))
