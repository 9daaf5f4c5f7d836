from typing import Optional, Tuple, Union
from torch import Tensor

OptTensor = Optional[Tensor]
PairTensor = Tuple[Tensor, Tensor]
PairOptTensor = Tuple[Optional[Tensor], Optional[Tensor]]
Adj = Union[Tensor, object]
Size = Optional[Tuple[int, int]]
