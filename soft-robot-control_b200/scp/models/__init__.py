from .ssm import SSMGuSTO          # noqa: F401
from .tpwl import TPWLGuSTO        # noqa: F401
