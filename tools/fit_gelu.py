"""Fit the tanh-form GELU used by the kernels (pcb_common.cuh PCB_GELU_C*) against the exact erf form."""
import numpy as np
from scipy.optimize import minimize
from scipy.special import erf

x = np.linspace(-8, 8, 160001)
phi = 0.5 * (1 + erf(x / np.sqrt(2)))
g, gp = x * phi, phi + x * np.exp(-0.5 * x * x) / np.sqrt(2 * np.pi)


def ev(c):
    t = np.tanh(x * (c[0] + c[1] * x ** 2 + c[2] * x ** 4))
    return 0.5 * x * (1 + t), 0.5 * (1 + t) + 0.5 * x * (1 - t * t) * (c[0] + 3 * c[1] * x ** 2 + 5 * c[2] * x ** 4)


def obj(c):
    a, b = ev(c)
    return np.abs(a - g).max() + 0.3 * np.abs(b - gp).max()


r = minimize(obj, [np.sqrt(2 / np.pi), np.sqrt(2 / np.pi) * 0.044715, 0.0], method="Nelder-Mead",
             options=dict(xatol=1e-10, fatol=1e-12, maxiter=20000))
a, b = ev(r.x)
print("coefficients", r.x, " max|GELU err| %.2e  max|GELU' err| %.2e" % (np.abs(a - g).max(), np.abs(b - gp).max()))
