"""Result / checkpoint file with the tree of the reference's ``InferenceFile``
(``gwin/io/hdf.py:101-132``; writers in ``gwin/sampler/base.py:152-186, 400-589`` and
``gwin/sampler/emcee.py:509-568``):

    attrs                      sampler, model, variable_params, sampling_params, niterations,
                               nwalkers, (ntemps), lognl, static params ...
    samples/{param}            [nwalkers, niterations]  or  [ntemps, nwalkers, niterations]
    model_stats/{stat}         same shapes
    acceptance_fraction        [nwalkers]  or  [ntemps, nwalkers]
    sampler_states/...         RNG state / counter, so that a run can resume bit-for-bit

h5py is not available in this environment, so the container is a NumPy ``.npz`` archive whose
member names are the HDF paths above; the layout, not the container, is what downstream readers
depend on.
"""
import json
import os

import numpy


class InferenceFile(object):
    name = "npz"
    samples_group = 'samples'
    stats_group = 'model_stats'
    sampler_group = 'sampler_states'

    def __init__(self, path, mode='r'):
        if mode not in ('r', 'w', 'a'):
            raise ValueError("mode must be 'r', 'w' or 'a'")
        self.path, self.mode = path, mode
        self.attrs, self._data = {}, {}
        if mode in ('r', 'a') and os.path.exists(path):
            with numpy.load(path, allow_pickle=False) as z:
                for k in z.files:
                    if k == '__attrs__':
                        self.attrs = json.loads(str(z[k]))
                    else:
                        self._data[k] = z[k]
        elif mode == 'r':
            raise IOError("no such file: %s" % path)

    # -- dict-like access to datasets by HDF-style path --------------------------
    def __getitem__(self, key):
        return self._data[key]

    def __setitem__(self, key, value):
        if self.mode == 'r':
            raise IOError("file is open read-only")
        self._data[key] = numpy.asarray(value)

    def __contains__(self, key):
        return key in self._data or any(k.startswith(key + '/') for k in self._data)

    def keys(self, group=None):
        if group is None:
            return sorted(self._data)
        pre = group + '/'
        return sorted(k[len(pre):] for k in self._data if k.startswith(pre))

    # -- properties of the reference class ------------------------------------------
    @property
    def sampler_name(self):
        return self.attrs["sampler"]

    @property
    def variable_params(self):
        return list(self.attrs["variable_params"])

    @property
    def niterations(self):
        return int(self.attrs["niterations"])

    @property
    def lognl(self):
        return self.attrs["lognl"]

    def append(self, key, value, axis=-1):
        """Extend a dataset along the iteration axis (the reference resizes the HDF dataset)."""
        value = numpy.asarray(value)
        if key in self._data:
            value = numpy.concatenate((self._data[key], value), axis=axis)
        self[key] = value

    def read_samples(self, parameters, thin_start=None, thin_interval=None, thin_end=None,
                     iteration=None, flatten=True, group=None):
        """Samples of the given parameters as a dict of arrays; thinning along the iteration
        axis as in ``BaseMCMCSampler.read_samples`` (reference sampler/base.py:645-735)."""
        group = group or self.samples_group
        if isinstance(parameters, str):
            parameters = [parameters]
        out = {}
        for p in parameters:
            arr = self._data['%s/%s' % (group, p)]
            if iteration is not None:
                arr = arr[..., iteration]
            else:
                arr = arr[..., slice(thin_start, thin_end, thin_interval)]
            out[p] = arr.reshape(-1) if flatten else arr
        return out

    def flush(self):
        if self.mode == 'r':
            return
        tmp = self.path + '.tmp.npz'
        numpy.savez(tmp, __attrs__=numpy.array(json.dumps(self.attrs, default=_jsonable)),
                    **self._data)
        os.replace(tmp, self.path)

    def close(self):
        self.flush()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _jsonable(x):
    if isinstance(x, (numpy.integer,)):
        return int(x)
    if isinstance(x, (numpy.floating,)):
        return float(x)
    if isinstance(x, numpy.ndarray):
        return x.tolist()
    return str(x)


def check_integrity(path):
    """Raises if the file cannot be read back or its datasets disagree on the number of
    iterations (reference ``io/hdf.py:755-801``)."""
    fp = InferenceFile(path, 'r')
    n = None
    for p in fp.keys(fp.samples_group):
        m = fp['%s/%s' % (fp.samples_group, p)].shape[-1]
        if n is not None and m != n:
            raise ValueError("parameters have different numbers of iterations")
        n = m
    if n is not None and n != fp.niterations:
        raise ValueError("niterations attribute (%d) does not match the samples (%d)"
                         % (fp.niterations, n))
    return True
