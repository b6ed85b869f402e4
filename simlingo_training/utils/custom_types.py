"""Wire types between the data layer, the CARLA agent and the model (reference
``simlingo_training/utils/custom_types.py:21-63``).  Field names *and order* are API: the agent builds
``DrivingInput(**dict)`` (``team_code/agent_simlingo.py:796``) and Lightning moves NamedTuples between
devices positionally."""
from typing import Dict, List, NamedTuple, Optional, Tuple

from torch import Tensor


class LanguageLabel(NamedTuple):
    phrase_ids: Tensor            # [B, L] int64 token ids (right/left padded)
    phrase_valid: Tensor          # [B, L] bool - fed to the model
    phrase_mask: Tensor           # [B, L] bool
    placeholder_values: list      # list[dict[token_id -> array[n, 2]]]
    language_string: list
    loss_masking: Tensor          # [B, L] bool - position takes part in the language loss


class DrivingInput(NamedTuple):
    camera_images: Tensor         # [B, T=1, NP, 3, 448, 448] pre-processed tiles
    image_sizes: Tensor
    camera_intrinsics: Tensor     # [B, N, 3, 3]
    camera_extrinsics: Tensor     # [B, N, 4, 4]
    vehicle_speed: Tensor         # [B, S] m/s
    target_point: Tensor          # [B, 2]
    prompt: LanguageLabel
    prompt_inference: LanguageLabel


class DrivingLabel(NamedTuple):
    waypoints: Tensor             # [B, F, 2]
    path: Tensor                  # [B, 20, 2]
    answer: LanguageLabel
    image_ff_org: Tensor
    eval_infos: Optional[Dict] = None


class DrivingExample(NamedTuple):
    driving_input: DrivingInput
    driving_label: DrivingLabel
    run_id: List[str]
    qa_templates: Optional[Tuple[str, str]] = None


class DrivingOutput(NamedTuple):
    waypoints: Tensor
    language_tokens: Tensor
    trajectory_tokens: Tensor


class TrainingOutput(NamedTuple):
    loss: Tensor                      # scalar
    loss_averages: Dict[str, Tensor]  # per-term scalar
    loss_values: Dict[str, Tensor]    # per-term [B]
    loss_counts: Dict[str, Tensor]    # per-term [B]
    driving_output: Optional[DrivingOutput] = None


class DatasetOutput(NamedTuple):
    """Produced by the (out-of-scope) dataset classes; kept so imports of this module keep working."""
    conversation: Optional[list]
    answer: Optional[str]
    image_ff: Optional[Tensor]
    image_ff_org_size: Optional[Tensor]
    waypoints: Optional[list]
    waypoints_1d: Optional[list]
    path: Optional[str]
    target_points: Optional[list]
    speed: Optional[float]
    placeholder_values: Optional[Dict]
    measurement_path: Optional[str]
    dataset: Optional[str]
    qa_templates: Optional[Tuple[str, str]] = None
    eval_infos: Optional[Dict] = None
