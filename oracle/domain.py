"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's `EvaluationDomain`.

Follows /root/reference/bellman/src/domain.rs:
  from_coeffs ......................... domain.rs:47-79
  fft / ifft .......................... domain.rs:81-99
  distribute_powers ................... domain.rs:101-113
  coset_fft / icoset_fft .............. domain.rs:115-125
  z / divide_by_z_on_coset ............ domain.rs:129-151
  mul_assign / sub_assign ............. domain.rs:154-189
  best_fft / serial_fft / parallel_fft  domain.rs:261-372

Only the `Scalar<S>` group is restated (the `Point<G>` path has no caller in the
crate).  Elements are canonical ints; generic over oracle.fields.PrimeField.
"""
from __future__ import annotations

from .multiexp import PolynomialDegreeTooLarge


def _bitreverse(n, l):
    r = 0
    for _ in range(l):
        r = (r << 1) | (n & 1)
        n >>= 1
    return r


def serial_fft(F, a, omega, log_n):
    """domain.rs:272-314 (in place on list `a`)."""
    n = len(a)
    assert n == 1 << log_n
    for k in range(n):
        rk = _bitreverse(k, log_n)
        if k < rk:
            a[rk], a[k] = a[k], a[rk]
    p = F.p
    m = 1
    for _ in range(log_n):
        w_m = pow(omega, n // (2 * m), p)
        k = 0
        while k < n:
            w = 1
            for j in range(m):
                t = a[k + j + m] * w % p
                a[k + j + m] = (a[k + j] - t) % p
                a[k + j] = (a[k + j] + t) % p
                w = w * w_m % p
            k += 2 * m
        m *= 2


def parallel_fft(F, a, omega, log_n, log_cpus):
    """domain.rs:316-372 (in place on list `a`)."""
    assert log_n >= log_cpus
    p = F.p
    num_cpus = 1 << log_cpus
    log_new_n = log_n - log_cpus
    tmp = [[0] * (1 << log_new_n) for _ in range(num_cpus)]
    new_omega = pow(omega, num_cpus, p)
    for j in range(num_cpus):
        omega_j = pow(omega, j, p)
        omega_step = pow(omega, j << log_new_n, p)
        elt = 1
        for i in range(1 << log_new_n):
            acc = 0
            for s in range(num_cpus):
                idx = (i + (s << log_new_n)) % (1 << log_n)
                acc = (acc + a[idx] * elt) % p
                elt = elt * omega_step % p
            tmp[j][i] = acc
            elt = elt * omega_j % p
        serial_fft(F, tmp[j], new_omega, log_new_n)
    mask = (1 << log_cpus) - 1
    for idx in range(len(a)):
        a[idx] = tmp[idx & mask][idx >> log_cpus]


def best_fft(F, a, omega, log_n, log_cpus=0):
    """domain.rs:261-269."""
    if log_n <= log_cpus:
        serial_fft(F, a, omega, log_n)
    else:
        parallel_fft(F, a, omega, log_n, log_cpus)


class EvaluationDomain:
    def __init__(self, F, coeffs, log_cpus=0):
        """from_coeffs, domain.rs:47-79."""
        self.F = F
        coeffs = list(coeffs)
        m, exp = 1, 0
        while m < len(coeffs):
            m *= 2
            exp += 1
            if exp >= F.S:
                raise PolynomialDegreeTooLarge()
        omega = F.root_of_unity
        for _ in range(exp, F.S):
            omega = omega * omega % F.p
        coeffs.extend([0] * (m - len(coeffs)))
        self.coeffs = coeffs
        self.exp = exp
        self.omega = omega
        self.omegainv = F.inv(omega)
        self.geninv = F.inv(F.generator)
        self.minv = F.inv(m % F.p)
        self.log_cpus = log_cpus

    def into_coeffs(self):
        return self.coeffs

    def fft(self):
        best_fft(self.F, self.coeffs, self.omega, self.exp, min(self.log_cpus, self.exp))

    def ifft(self):
        best_fft(self.F, self.coeffs, self.omegainv, self.exp, min(self.log_cpus, self.exp))
        p, minv = self.F.p, self.minv
        self.coeffs = [v * minv % p for v in self.coeffs]

    def distribute_powers(self, g):
        p = self.F.p
        u = 1
        out = []
        for v in self.coeffs:
            out.append(v * u % p)
            u = u * g % p
        self.coeffs = out

    def coset_fft(self):
        self.distribute_powers(self.F.generator)
        self.fft()

    def icoset_fft(self):
        self.ifft()
        self.distribute_powers(self.geninv)

    def z(self, tau):
        return (pow(tau, len(self.coeffs), self.F.p) - 1) % self.F.p

    def divide_by_z_on_coset(self):
        i = self.F.inv(self.z(self.F.generator))
        p = self.F.p
        self.coeffs = [v * i % p for v in self.coeffs]

    def mul_assign(self, other):
        assert len(self.coeffs) == len(other.coeffs)
        p = self.F.p
        self.coeffs = [a * b % p for a, b in zip(self.coeffs, other.coeffs)]

    def sub_assign(self, other):
        assert len(self.coeffs) == len(other.coeffs)
        p = self.F.p
        self.coeffs = [(a - b) % p for a, b in zip(self.coeffs, other.coeffs)]


def dft_by_definition(F, a, omega):
    """out[k] = sum_j a[j] * omega^(jk) -- the mathematical object every fft variant
    must equal (natural order in, natural order out)."""
    n, p = len(a), F.p
    out = []
    for k in range(n):
        wk = pow(omega, k, p)
        acc, w = 0, 1
        for j in range(n):
            acc = (acc + a[j] * w) % p
            w = w * wk % p
        out.append(acc)
    return out
