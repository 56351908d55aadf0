from .serial_chain import SerialChainFK, PandaFK, PANDA_CHAIN  # noqa: F401
