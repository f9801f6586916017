from heybuddy_b200.util.log_util import logger
from heybuddy_b200.util.native_model import PretrainedNativeModel, PretrainedONNXModel
from heybuddy_b200.util.audio_util import audio_to_bct_tensor
from heybuddy_b200.util.string_util import safe_name

__all__ = ["logger", "PretrainedNativeModel", "PretrainedONNXModel", "audio_to_bct_tensor", "safe_name"]
