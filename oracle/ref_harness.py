"""TEST INFRASTRUCTURE ONLY.  Runs the UNMODIFIED reference (nv-tlabs/ASE @ /root/reference) on CPU
torch through the stand-in packages in oracle/shims (isaacgym, rl_games, gym, tensorboardX).
Only usable in the build container (the GPU box has no /root/reference): it is used by
oracle/gen_golden.py to write tests/golden/*.pt and by tests/test_oracle_vs_reference.py to pin
oracle/ase_oracle.py (the travelling CPU restatement) to the reference's own code."""
import os
import sys
import copy

import numpy as np
import torch
import yaml

REF_ROOT = os.environ.get('ASE_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'ase', 'learning'))


def _setup_path():
    for p in (_SHIMS, os.path.join(REF_ROOT, 'ase')):
        if p not in sys.path:
            sys.path.insert(0, p)


def import_env_fns():
    """-> (humanoid module, humanoid_amp module, utils.torch_utils) of the reference."""
    _setup_path()
    from env.tasks import humanoid, humanoid_amp
    from utils import torch_utils
    return humanoid, humanoid_amp, torch_utils


class _Task:
    def __init__(self, num_envs):
        self.num_envs = num_envs
        self.progress_buf = torch.zeros(num_envs, dtype=torch.long)
        self.viewer = None


class _Env:
    def __init__(self, num_envs, obs_dim, amp_dim, act_dim):
        from gym import spaces
        self.task = _Task(num_envs)
        self.observation_space = spaces.Box(np.ones(obs_dim) * -np.inf, np.ones(obs_dim) * np.inf)
        self.amp_observation_space = spaces.Box(np.ones(amp_dim) * -np.inf, np.ones(amp_dim) * np.inf)
        self.action_space = spaces.Box(np.ones(act_dim) * -1., np.ones(act_dim) * 1.)
        self.num_states = 0
        self._demo_fn = None

    def fetch_amp_obs_demo(self, n):
        return self._demo_fn(n)


class FakeVecEnv:
    """What run.py:100-145 (RLGPUEnv) exposes to the agents, with scripted tensors instead of a sim.
    `script` is a callable(step_index, actions) -> (obs, rew, dones, {'amp_obs','terminate'})."""

    def __init__(self, num_envs, obs_dim=253, amp_dim=1400, act_dim=31):
        self.env = _Env(num_envs, obs_dim, amp_dim, act_dim)
        self.script = None
        self.reset_fn = None
        self.t = 0

    def get_env_info(self):
        return {'action_space': self.env.action_space,
                'observation_space': self.env.observation_space,
                'amp_observation_space': self.env.amp_observation_space}

    def step(self, actions):
        out = self.script(self.t, actions)
        self.t += 1
        self.env.task.progress_buf += 1
        return out

    def reset(self, env_ids=None):
        return self.reset_fn(env_ids)

    def set_env_state(self, s):
        pass


def load_train_cfg(name):
    with open(os.path.join(REF_ROOT, 'ase', 'data', 'cfg', 'train', 'rlg', name)) as f:
        return yaml.load(f, Loader=yaml.SafeLoader)


def make_ref_agent(kind='ase', num_envs=8, overrides=None, obs_dim=253, amp_dim=1400, act_dim=31, seed=0):
    """Instantiate the reference's own ASEAgent / AMPAgent / HRL-style CommonAgent on CPU.
    Mirrors what rl_games' Runner.load()+algo_factory.create() hand to the agent (run.py:153-170)."""
    _setup_path()
    from rl_games.common.tr_helpers import DefaultRewardsShaper
    from learning import amp_agent, ase_agent, common_agent
    from learning import amp_models, ase_models, amp_network_builder, ase_network_builder
    torch.manual_seed(seed)
    np.random.seed(seed)
    yaml_name = {'ase': 'ase_humanoid.yaml', 'amp': 'amp_humanoid.yaml', 'hrl': 'hrl_humanoid.yaml'}[kind]
    params = copy.deepcopy(load_train_cfg(yaml_name))['params']
    config = params['config']
    config.update(overrides or {})
    if kind == 'ase':
        builder = ase_network_builder.ASEBuilder()
        builder.load(params['network'])
        config['network'] = ase_models.ModelASEContinuous(builder)
        cls = ase_agent.ASEAgent
    elif kind == 'hrl':
        # the HLC learner of HRLAgent is CommonAgent.calc_gradients over HRLBuilder's network (hrl_agent.py:25, hrl_network_builder.py:8-39)
        from learning import hrl_models, hrl_network_builder
        builder = hrl_network_builder.HRLBuilder()
        builder.load(params['network'])
        config['network'] = hrl_models.ModelHRLContinuous(builder)
        cls = common_agent.CommonAgent
    else:
        builder = amp_network_builder.AMPBuilder()
        builder.load(params['network'])
        config['network'] = amp_models.ModelAMPContinuous(builder)
        cls = amp_agent.AMPAgent
    config['reward_shaper'] = DefaultRewardsShaper(**config['reward_shaper'])
    config['num_actors'] = num_envs
    config['device'] = 'cpu'
    vec_env = FakeVecEnv(num_envs, obs_dim, amp_dim, act_dim)
    config['env_info'] = vec_env.get_env_info()
    agent = cls(base_name='oracle', config=config)
    agent.vec_env = vec_env
    return agent, params
