# Regular package whose search path is EXTENDED with the reference's openfold/ directory, so that
# openfold.utils.rigid_utils and openfold.model.structure_module resolve here while openfold.np, openfold.data,
# openfold.utils.tensor_utils, ... still come from the reference.
import pkgutil
__path__ = pkgutil.extend_path(__path__, __name__)
