"""[recollection of rl_games 1.1.4 common/experience.py: ExperienceBuffer] (continuous-action subset)."""
import numpy as np
import torch
from rl_games.algos_torch.torch_ext import numpy_to_torch_dtype_dict


class ExperienceBuffer:
    def __init__(self, env_info, algo_info, device, aux_tensor_dict=None):
        self.env_info = env_info
        self.algo_info = algo_info
        self.device = device
        self.num_agents = env_info.get('agents', 1)
        self.action_space = env_info['action_space']
        self.num_actors = algo_info['num_actors']
        self.horizon_length = algo_info['horizon_length']
        self.has_central_value = algo_info['has_central_value']
        self.use_action_masks = algo_info.get('use_action_masks', False)
        batch_size = self.num_actors * self.num_agents
        self.is_discrete = False
        self.is_multi_discrete = False
        self.is_continuous = True
        self.obs_base_shape = (self.horizon_length, self.num_agents * self.num_actors)
        self.state_base_shape = (self.horizon_length, self.num_actors)
        self.actions_shape = (self.action_space.shape[0],)
        self.actions_num = self.action_space.shape[0]
        self.tensor_dict = {}
        self._init_from_env_info(self.env_info)
        self.aux_tensor_dict = aux_tensor_dict

    def _init_from_env_info(self, env_info):
        obs_base_shape = self.obs_base_shape
        self.tensor_dict['obses'] = self._create_tensor_from_space(env_info['observation_space'], obs_base_shape)
        val_space = _Box((env_info.get('value_size', 1),), np.float32)
        self.tensor_dict['rewards'] = self._create_tensor_from_space(val_space, obs_base_shape)
        self.tensor_dict['values'] = self._create_tensor_from_space(val_space, obs_base_shape)
        self.tensor_dict['neglogpacs'] = self._create_tensor_from_space(_Box((), np.float32), obs_base_shape)
        self.tensor_dict['dones'] = self._create_tensor_from_space(_Box((), np.uint8), obs_base_shape)
        self.tensor_dict['actions'] = self._create_tensor_from_space(_Box(self.actions_shape, np.float32), obs_base_shape)
        self.tensor_dict['mus'] = self._create_tensor_from_space(_Box(self.actions_shape, np.float32), obs_base_shape)
        self.tensor_dict['sigmas'] = self._create_tensor_from_space(_Box(self.actions_shape, np.float32), obs_base_shape)

    def _create_tensor_from_space(self, space, base_shape):
        dtype = numpy_to_torch_dtype_dict[np.dtype(space.dtype)]
        return torch.zeros(base_shape + tuple(space.shape), dtype=dtype, device=self.device)

    def update_data(self, name, index, val):
        if type(val) is dict:
            for k, v in val.items():
                self.tensor_dict[name][k][index, :] = v
        else:
            self.tensor_dict[name][index, :] = val

    def get_transformed(self, transform_op):
        res_dict = {}
        for k, v in self.tensor_dict.items():
            res_dict[k] = transform_op(v)
        return res_dict

    def get_transformed_list(self, transform_op, tensor_list):
        res_dict = {}
        for k in tensor_list:
            v = self.tensor_dict.get(k)
            if v is None:
                continue
            res_dict[k] = transform_op(v)
        return res_dict


class _Box:
    def __init__(self, shape, dtype):
        self.shape = shape
        self.dtype = dtype
