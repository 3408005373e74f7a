"""``Agent``: host-side record of one agent with the reference's constructor and attribute names
(reference ``envs/agent.py:29-138``).  The per-step update of these attributes happens on the device; the
environment facade copies the device columns back into the Agent objects after every step so that code written
against ``env.agents[i].pos_global_frame`` etc. keeps working.
"""
import math

import numpy as np

from . import Config as _DefaultConfig  # noqa: F401  (import-time singleton, like the reference)


def _cfg():
    from . import Config
    return Config


class Agent(object):
    def __init__(self, start_x, start_y, goal_x, goal_y, radius, pref_speed, initial_heading, policy, dynamics_model,
                 sensors, id):
        self.policy = policy()
        self.dynamics_model = dynamics_model(self)
        self.sensors = [sensor() for sensor in sensors]
        self.chosen_action_dict = {}
        self.num_actions_to_store = 2
        self.action_dim = 2
        self.id = id
        self.dist_to_goal = 0.0
        C = _cfg()
        self.near_goal_threshold = C.NEAR_GOAL_THRESHOLD
        self.dt_nominal = C.DT
        self.num_other_agents_observed = 0
        self.min_x, self.max_x, self.min_y, self.max_y = -20.0, 20.0, -20.0, 20.0
        self.t_offset = None
        self.global_state_dim = 11
        self.ego_state_dim = 3
        self.max_heading_change = np.pi / 3
        self.max_speed = 1.0
        self.sensor_data = {}
        self.reset(px=start_x, py=start_y, gx=goal_x, gy=goal_y, pref_speed=pref_speed, radius=radius,
                   heading=initial_heading)

    def reset(self, px=None, py=None, gx=None, gy=None, pref_speed=None, radius=None, heading=None):
        """Reference ``Agent.reset`` (``agent.py:59-138``)."""
        C = _cfg()
        if px is not None and py is not None:
            self.pos_global_frame = np.array([px, py], dtype="float64")
        if gx is not None and gy is not None:
            self.goal_global_frame = np.array([gx, gy], dtype="float64")
        self.vel_global_frame = np.array([0.0, 0.0], dtype="float64")
        self.speed_global_frame = 0.0
        if heading is None:
            v = self.goal_global_frame - self.pos_global_frame
            self.heading_global_frame = np.arctan2(v[1], v[0])
        else:
            self.heading_global_frame = np.float64(heading)
        self.delta_heading_global_frame = 0.0
        self.speed_ego_frame = 0.0
        self.heading_ego_frame = 0.0
        self.vel_ego_frame = np.array([0.0, 0.0])
        self.past_actions = np.zeros((self.num_actions_to_store, self.action_dim))
        if radius is not None:
            self.radius = radius
        if pref_speed is not None:
            self.pref_speed = pref_speed
        d = self.pos_global_frame - self.goal_global_frame
        self.straight_line_time_to_reach_goal = (math.sqrt(d[0] * d[0] + d[1] * d[1])
                                                 - self.near_goal_threshold) / self.pref_speed
        self.time_remaining_to_reach_goal = max(C.MAX_TIME_RATIO * self.straight_line_time_to_reach_goal,
                                                self.dt_nominal)
        self.t = 0.0
        self.step_num = 0
        self.is_at_goal = False
        self.was_at_goal_already = False
        self.was_in_collision_already = False
        self.in_collision = False
        self.ran_out_of_time = False
        self.num_states_in_history = int(1.2 * self.time_remaining_to_reach_goal / self.dt_nominal)
        self.global_state_history = np.empty((self.num_states_in_history, self.global_state_dim))
        self.ego_state_history = np.empty((self.num_states_in_history, self.ego_state_dim))
        self.past_global_velocities = np.zeros((self.num_actions_to_store, 2))
        self.other_agent_states = np.zeros((7,))
        self.dynamics_model.update_ego_frame()
        self.min_dist_to_other_agents = np.inf
        self.turning_dir = 0.0
        self.is_done = False

    def __deepcopy__(self, memo):
        cls = self.__class__
        obj = cls.__new__(cls)
        for k, v in self.__dict__.items():
            if k != "policy":
                setattr(obj, k, v)
        return obj

    def get_ref(self):
        """Reference ``agent.py:329-349``."""
        goal_direction = self.goal_global_frame - self.pos_global_frame
        self.dist_to_goal = math.sqrt(goal_direction[0] ** 2 + goal_direction[1] ** 2)
        ref_prll = goal_direction / self.dist_to_goal if self.dist_to_goal > 1e-8 else goal_direction
        ref_orth = np.array([-ref_prll[1], ref_prll[0]])
        return ref_prll, ref_orth

    def set_state(self, px, py, vx=None, vy=None, heading=None):
        """Reference ``agent.py:155-190`` (for ExternalDynamics agents); the facade uploads the new state."""
        if vx is None or vy is None:
            if self.step_num == 0:
                self.vel_global_frame = np.array([0.0, 0.0])
            else:
                self.vel_global_frame = (np.array([px, py]) - self.pos_global_frame) / self.dt_nominal
        else:
            self.vel_global_frame = np.array([vx, vy], dtype="float64")
        if heading is None:
            heading = np.arctan2(self.vel_global_frame[1], self.vel_global_frame[0])
        d = heading - self.heading_global_frame
        while d >= np.pi:
            d -= 2 * np.pi
        while d < -np.pi:
            d += 2 * np.pi
        self.delta_heading_global_frame = d
        self.pos_global_frame = np.array([px, py], dtype="float64")
        self.speed_global_frame = float(np.linalg.norm(self.vel_global_frame))
        self.heading_global_frame = np.float64(heading)
        self._state_dirty = True

    def to_vector(self):
        global_state = np.array([self.t, self.pos_global_frame[0], self.pos_global_frame[1],
                                 self.goal_global_frame[0], self.goal_global_frame[1], self.radius, self.pref_speed,
                                 self.vel_global_frame[0], self.vel_global_frame[1], self.speed_global_frame,
                                 self.heading_global_frame])
        ego_state = np.array([self.t, self.dist_to_goal, self.heading_ego_frame])
        return global_state, ego_state

    def get_sensor_data(self, sensor_name):
        if sensor_name in self.sensor_data:
            return self.sensor_data[sensor_name]

    def get_agent_data(self, attribute):
        return getattr(self, attribute)

    def get_agent_data_equiv(self, attribute, value):
        obj = self
        for part in attribute.split("."):
            obj = getattr(obj, part)
        return obj == value

    def get_observation_dict(self, agents=None):
        """Reference ``agent.py:323-327`` (host view of the observation the device packed)."""
        C = _cfg()
        return {s: np.array(self._obs_value(s)) for s in C.STATES_IN_OBS}

    def _obs_value(self, state):
        table = {"dist_to_goal": lambda: self.dist_to_goal, "radius": lambda: self.radius,
                 "heading_ego_frame": lambda: self.heading_ego_frame, "pref_speed": lambda: self.pref_speed,
                 "num_other_agents": lambda: self.num_other_agents_observed,
                 "other_agent_states": lambda: self.other_agent_states,
                 "other_agents_states": lambda: self.sensor_data.get("other_agents_states"),
                 "laserscan": lambda: self.sensor_data.get("laserscan"),
                 "is_learning": lambda: self.policy.str == "learning"}
        return table[state]()

    def print_agent_info(self):
        print("----------\nGlobal Frame:\n(px,py):", self.pos_global_frame, "\n(vx,vy):", self.vel_global_frame,
              "\nspeed:", self.speed_global_frame, "\nheading:", self.heading_global_frame, "\nBody Frame:\n(vx,vy):",
              self.vel_ego_frame, "\nheading:", self.heading_ego_frame, "\n----------")
