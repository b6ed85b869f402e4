"""Drop-in mirror of the reference's ``simlingo_training`` package for the VLA hot path only
(``simlingo_training.models.*`` and ``simlingo_training.utils.custom_types``).  Same module paths, class
names, constructor / forward signatures and ``state_dict`` keys as the reference, so that
``team_code/agent_simlingo.py`` and the Lightning training loop can ``hydra.utils.instantiate`` the
``_target_`` strings of ``simlingo_training/config.py:46,71,104`` unchanged - with every heavy op routed to
the sm_100a kernels in ``simlingo_b200``."""
