from .nn_ops import NN_Trainer, accuracy, svd_encode
from .master import SyncReplicasMaster_NN, GradientAccumulator, build_coder, STEP_START_
from .worker import DistributedWorker
from .evaluator import DistributedEvaluator
from .flat import FlatLayout, bind_parameters, bind_gradients
