from .acquisition import Acquisition, AcquisitionAggregation, AcquisitionProduct, AcquisitionSum, MCMCAcquistion
from .ei import ExpectedImprovement
from .hvpoi import HVProbabilityOfImprovement
from .lcb import LowerConfidenceBound
from .pof import ProbabilityOfFeasibility
from .poi import ProbabilityOfImprovement
from .mes import MinValueEntropySearch
