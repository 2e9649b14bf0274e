"""Host orchestration of PolyModel.fit on the device (filled in with the Gram / Cholesky kernels)."""


def fit_polymodel(model, x, y, logp, w, comm=None, refine=1):
    raise NotImplementedError('fit kernels are not built yet')


def set_bound(model, x, logp):
    raise NotImplementedError('fit kernels are not built yet')


def ellipsoid(model, x, alpha_p):
    raise NotImplementedError('fit kernels are not built yet')
