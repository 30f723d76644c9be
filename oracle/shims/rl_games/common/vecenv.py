vecenv_config = {}
def register(config_name, func): vecenv_config[config_name] = func
def create_vec_env(config_name, num_actors, **kwargs):
    return vecenv_config[config_name](config_name, num_actors, **kwargs)
class IVecEnv: pass
