"""Error types of the environment API.  ``InvalidAgentError`` mirrors the reference's only error path
(lib/utils/exceptions.py:6-7, raised by AirHockey.update_state for an unknown agent_name, airhockey.py:108-110)."""


class InvalidAgentError(Exception):
    pass
