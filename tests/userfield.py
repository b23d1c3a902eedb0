"""The user-defined field of examples/Creating new fields.ipynb cell 10 (ChargedDipole) as a
rapt_b200 field plugin: host-side B/E as in the notebook + the CUDA snippet NVRTC compiles."""
import numpy as np


def make_charged_dipole():
    from rapt_b200 import fields

    class ChargedDipole(fields._Field):
        cuda_has_E = True
        cuda_source = r'''
__device__ void rapt_user_B(double t, double x, double y, double z, const double* prm, double* B) {
    double p = pow(x*x + y*y + z*z, 5.0/2.0);
    B[0] = prm[0] * (3*x*z) / p; B[1] = prm[0] * (3*y*z) / p; B[2] = prm[0] * (2*z*z - x*x - y*y) / p;
}
__device__ void rapt_user_E(double t, double x, double y, double z, const double* prm, double* E) {
    double p = pow(x*x + y*y + z*z, 3.0/2.0), kq = prm[2] * prm[1];
    E[0] = kq * x / p; E[1] = kq * y / p; E[2] = kq * z / p;
}
'''

        def __init__(self, B0=1, Q=1):
            fields._Field.__init__(self)
            self.B0 = B0; self.Q = Q; self._k = 8.9875517873681764e9
            self.static = False

        def cuda_params(self):
            return [self.B0, self.Q, self._k]

        def B(self, tpos):
            t, x, y, z = tpos
            return self.B0 * np.array([3*x*z, 3*y*z, (2*z*z - x*x - y*y)]) / pow(x*x + y*y + z*z, 5.0/2.0)

        def E(self, tpos):
            t, x, y, z = tpos
            return self._k * self.Q * np.array([x, y, z]) / pow(x*x + y*y + z*z, 3.0/2.0)

    return ChargedDipole
