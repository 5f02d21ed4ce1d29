"""Drop-in for the reference's ``dirichlet/psignn/model.py``: same class names, constructor signatures,
``state_dict`` keys and return values; the implicit solve runs on the sm_100a kernels."""
from ...model import (MLP, Phi_to, Phi_from, Encoder, Decoder, Autoencoder, DeepEquilibrium,  # noqa: F401
                      initialize_weights_xavier, jac_loss_estimate, power_method)
from ...model import FunctionDirichlet as Function              # noqa: F401
from ...model import ModelDEQDSSDirichlet as ModelDEQDSS        # noqa: F401
# evaluation-form variants of the reference's tests/model_psignn.py (solver dict out of the DEQ wrapper, `nsteps` in the loss dict)
from ...model import ModelPSIGNN, ModelPSIGNNIterative          # noqa: F401,E402
