"""ORACLE SUPPORT: imported, never called (cfg.gnet.imfeats is False)."""
