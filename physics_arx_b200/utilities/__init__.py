"""``tigre.utilities`` surface used by the reference: ``CTnoise`` (:88) and ``gpu`` (:24,31)."""
from . import CTnoise, gpu  # noqa: F401
