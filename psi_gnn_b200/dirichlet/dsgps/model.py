"""Drop-in for the reference's ``dirichlet/dsgps/model.py`` (inference on the shared fused layer kernel)."""
from ...baselines import ModelDSGPS, MLPActivation                                       # noqa: F401
from ...model import MLP, Phi_to, Phi_from, Encoder, Decoder, Autoencoder, initialize_weights_xavier   # noqa: F401
