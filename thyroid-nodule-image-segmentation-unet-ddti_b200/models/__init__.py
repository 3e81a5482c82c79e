"""Drop-in module namespace: same import names as the reference's models/ package (model.py, loss.py, vnet.py)."""
