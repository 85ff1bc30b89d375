"""Training path (train-mode BatchNorm forward, backward, Adam) — see DESIGN.md."""
from . import _lib as L


def forward_train(model, x, t, c, shortcut):
    raise L.CdmError("train-mode forward is not built yet in this revision: call model.eval() "
                     "(sampling / likelihood / ELBO paths are available)")
