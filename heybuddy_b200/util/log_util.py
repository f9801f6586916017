"""Logger shared by the package (reference: util/log_util.py:40-47 uses the name "heybuddy")."""
import logging

logger = logging.getLogger("heybuddy")
