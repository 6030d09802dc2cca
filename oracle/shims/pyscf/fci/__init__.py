from . import addons, selected_ci  # noqa: F401
