"""Empty stand-in (oracle shim): the hot path never touches gymapi."""
class Vec3:
    def __init__(self, x=0.0, y=0.0, z=0.0):
        self.x, self.y, self.z = x, y, z
class PlaneParams: pass
class Transform: pass
SIM_PHYSX = 1
UP_AXIS_Z = 1
DOF_MODE_POS = 1
DOF_MODE_EFFORT = 2
ENV_SPACE = 0
KEY_ESCAPE = KEY_V = KEY_Q = 0
