from .linear import (MinMaxQuantLinear, PTQSLQuantLinear, PostGeluPTQSLQuantLinear, PTQSLBatchingQuantLinear,
                     PostGeluPTQSLBatchingQuantLinear)
from .matmul import (MinMaxQuantMatMul, PTQSLQuantMatMul, SoSPTQSLQuantMatMul, PTQSLBatchingQuantMatMul,
                     SoSPTQSLBatchingQuantMatMul)
from .conv import MinMaxQuantConv2d, PTQSLQuantConv2d, ChannelwiseBatchingQuantConv2d
