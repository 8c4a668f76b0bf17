from .linear import (MinMaxQuantLinear, PTQSLQuantLinear, PostGeluPTQSLQuantLinear, PTQSLBatchingQuantLinear,
                     PostGeluPTQSLBatchingQuantLinear)
from .matmul import (MinMaxQuantMatMul, PTQSLQuantMatMul, SoSPTQSLQuantMatMul, PTQSLBatchingQuantMatMul,
                     SoSPTQSLBatchingQuantMatMul)
