"""Reader for the reference's saved models (tf.train.Saver bundles written by GANMF.saveModel,
GANRec/GANMF.py:309-314): `<name>.data-00000-of-00001` holds the raw little-endian fp32 tensors
concatenated in alphabetical order of their TF variable names (verified on the one surviving
reference checkpoint), so the shapes from build_params.pkl + the URM are enough."""
import numpy as np


def read_tf_bundle(data_path, shapes):
    raw = np.fromfile(data_path, dtype="<f4")
    out, off = {}, 0
    for name in sorted(shapes):
        n = int(np.prod(shapes[name]))
        out[name] = raw[off:off + n].reshape(shapes[name]).copy()
        off += n
    if off != raw.size:
        raise ValueError("TF bundle %s holds %d floats, the model needs %d" % (data_path, raw.size, off))
    return out
