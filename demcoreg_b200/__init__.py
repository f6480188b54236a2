"""demcoreg_b200: B200-native co-registration hot path of dshean/demcoreg."""
