"""Host runtime of the B200-native ATM-VFI forward."""
