"""Drop-in for the reference's ``mixed/psignn/model.py`` (Dirichlet + homogeneous Neumann boundary conditions)."""
from ...model import (MLP, Phi_to, Phi_from, Encoder, Decoder, Autoencoder, DeepEquilibrium,  # noqa: F401
                      initialize_weights_xavier, jac_loss_estimate, power_method)
from ...model import FunctionMixed as Function                  # noqa: F401
from ...model import ModelDEQDSSMixed as ModelDEQDSS            # noqa: F401
